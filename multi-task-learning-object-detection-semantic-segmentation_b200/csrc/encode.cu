// SSD target encode (reference: datacoder.py:205-300).
//
// Parity contract: IoU is evaluated with exactly the reference's fp32 expression tree
// (one rounding per op, IEEE division, association (area_a + area_g) - inter); both arg-max
// reductions resolve ties to the lowest index; the concat -> UniqueV2 -> sequential scatter of
// :244/:286 is replaced by its closed form:
//     S1(a) = { g : argmax_a iou[:,g] == a  and  max_a iou[:,g] > 0 }
//     g2(a) = argmax_g iou[a,:]  if  max_g iou[a,:] > thr
//     match(a) = g2 if g2 exists and g2 not in S1(a);  else max(S1(a)) if non-empty;  else background.
//
// Two kernels, no A x G matrix is ever materialised:
//   match : grid (anchor tiles, images).  GT boxes of the image sit in shared memory, each thread keeps
//           kApt anchors in registers, per-anchor running arg-max in registers, per-GT arg-max by warp
//           REDUX (max of the IoU bit pattern, then min anchor index among the tied lanes) into a
//           warp-private shared row, one 64-bit atomicMax per (CTA, GT) into the workspace.
//           Pairs with empty intersection (78% on the reference anchor grid) skip the division: their
//           IoU is exactly +0 and can never win a strict '>' comparison.
//   emit  : per anchor, resolves the closed form from the two workspace arrays and writes the one-hot
//           labels and the offsets with coalesced 128-bit stores.
#include "common.cuh"

namespace {

constexpr int kMatchThreads = 256;
constexpr int kApt = 2;          // anchors per thread in the match kernel
constexpr int kGtTile = 256;     // GT boxes staged per shared-memory tile
constexpr int kEmitThreads = 256;

struct GtBox {
  float x0, y0, x1, y1;
};

// packed (iou bits, inverted anchor index): larger = better IoU, ties -> lower anchor index
__device__ __forceinline__ unsigned long long pack_best(float iou, int a) {
  return ((unsigned long long)__float_as_uint(iou) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)a);
}

__global__ void __launch_bounds__(kMatchThreads)
    ssdx_encode_match_kernel(const float4* __restrict__ anc_corners, const float* __restrict__ anc_area,
                             const float4* __restrict__ gt_xyxy, const int* __restrict__ gt_count, const int A,
                             const int Gmax, const float iou_thr, unsigned long long* __restrict__ gt_best,
                             int* __restrict__ anchor_g2) {
  __shared__ float4 s_box[kGtTile];
  __shared__ float s_area[kGtTile];
  __shared__ unsigned long long s_wbest[(kMatchThreads / 32)][kGtTile];  // warp-private running best per GT

  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);
  const int a_base = blockIdx.x * (kMatchThreads * kApt);

  float4 an[kApt];
  float aa[kApt];
  int aidx[kApt];
  float best_v[kApt];
  int best_g[kApt];
#pragma unroll
  for (int u = 0; u < kApt; ++u) {
    // warp w owns 32*kApt consecutive anchors; lane l owns a_base + w*32*kApt + u*32 + l (ascending in u)
    aidx[u] = a_base + warp * (32 * kApt) + u * 32 + lane;
    const bool ok = aidx[u] < A;
    an[u] = ok ? __ldg(anc_corners + aidx[u]) : make_float4(0.f, 0.f, -2.f, -2.f);  // empty box: inter == 0
    aa[u] = ok ? __ldg(anc_area + aidx[u]) : 1.0f;
    best_v[u] = 0.0f;  // np.argmax over all-zero row -> index 0
    best_g[u] = 0;
  }

  for (int g0 = 0; g0 < G; g0 += kGtTile) {
    const int gn = min(kGtTile, G - g0);
    __syncthreads();
    for (int g = tid; g < gn; g += kMatchThreads) {
      const float4 bx = __ldg(gt_xyxy + (size_t)b * Gmax + g0 + g);
      s_box[g] = bx;
      // datacoder.py:206  (xmax-xmin+1)*(ymax-ymin+1)
      s_area[g] = __fmul_rn(__fadd_rn(__fsub_rn(bx.z, bx.x), 1.0f), __fadd_rn(__fsub_rn(bx.w, bx.y), 1.0f));
    }
    __syncthreads();

    for (int g = 0; g < gn; ++g) {
      const float4 gb = s_box[g];
      const float ga = s_area[g];
      float mv = 0.0f;
      int ma = 0x7FFFFFFF;
#pragma unroll
      for (int u = 0; u < kApt; ++u) {
        // datacoder.py:210-216
        const float iw = __fadd_rn(__fsub_rn(fminf(an[u].z, gb.z), fmaxf(an[u].x, gb.x)), 1.0f);
        const float ih = __fadd_rn(__fsub_rn(fminf(an[u].w, gb.w), fmaxf(an[u].y, gb.y)), 1.0f);
        const float inter = __fmul_rn(fmaxf(0.0f, iw), fmaxf(0.0f, ih));
        if (inter > 0.0f) {
          // datacoder.py:220  inter / ((area_a + area_g) - inter), IEEE division
          const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa[u], ga), inter));
          if (iou > best_v[u]) {  // strict: ties keep the lower GT index
            best_v[u] = iou;
            best_g[u] = g0 + g;
          }
          if (iou > mv) {  // strict: ties keep the lower anchor index (u ascending)
            mv = iou;
            ma = aidx[u];
          }
        }
      }
      // warp arg-max for this GT: IoU >= 0 so the bit pattern is monotone
      const unsigned mbits = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(mv));
      unsigned long long wb = 0ull;
      if (mbits != 0u) {  // warp-uniform
        // lowest anchor index among the lanes that tie on the maximum (every lane takes part in the REDUX)
        const int amin = __reduce_min_sync(0xFFFFFFFFu, (__float_as_uint(mv) == mbits) ? ma : 0x7FFFFFFF);
        wb = pack_best(__uint_as_float(mbits), amin);
      }
      if (lane == 0) s_wbest[warp][g] = wb;  // each (warp, GT) is visited exactly once per tile
    }
    __syncthreads();
    // one global atomic per (CTA, GT) that saw a positive IoU
    for (int g = tid; g < gn; g += kMatchThreads) {
      unsigned long long m = 0ull;
#pragma unroll
      for (int w = 0; w < kMatchThreads / 32; ++w) {
        const unsigned long long v = s_wbest[w][g];
        m = v > m ? v : m;
      }
      if (m != 0ull) atomicMax(gt_best + (size_t)b * Gmax + g0 + g, m);
    }
  }

#pragma unroll
  for (int u = 0; u < kApt; ++u)
    if (aidx[u] < A) anchor_g2[(size_t)b * A + aidx[u]] = (G > 0 && best_v[u] > iou_thr) ? best_g[u] : -1;
}

__global__ void __launch_bounds__(kEmitThreads)
    ssdx_encode_emit_kernel(const float4* __restrict__ anc_cent, const float4* __restrict__ gt_xyxy,
                            const int* __restrict__ gt_label, const int* __restrict__ gt_count, const int A,
                            const int Gmax, const int C, const float sx, const float sy, const float sw,
                            const float sh, const unsigned long long* __restrict__ gt_best,
                            const int* __restrict__ anchor_g2, float* __restrict__ out_labels,
                            float4* __restrict__ out_offsets, int* __restrict__ out_match) {
  __shared__ int s_s1max[kEmitThreads];
  __shared__ int s_lab[kEmitThreads];
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int a_base = blockIdx.x * kEmitThreads;
  const int a = a_base + tid;
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);
  const unsigned long long* best = gt_best + (size_t)b * Gmax;

  s_s1max[tid] = -1;
  __syncthreads();
  // S1: GT boxes whose best anchor (IoU > 0) lies in this tile
  for (int g = tid; g < G; g += kEmitThreads) {
    const unsigned long long p = best[g];
    if ((p >> 32) != 0ull) {
      const int ba = (int)(0xFFFFFFFFu - (unsigned)(p & 0xFFFFFFFFull));
      if (ba >= a_base && ba < a_base + kEmitThreads) atomicMax(&s_s1max[ba - a_base], g);
    }
  }
  __syncthreads();

  int match = -1;
  if (a < A) {
    const int g2 = anchor_g2[(size_t)b * A + a];
    const int s1 = s_s1max[tid];
    if (g2 >= 0) {
      const unsigned long long p = best[g2];
      const bool in_s1 = ((p >> 32) != 0ull) && ((int)(0xFFFFFFFFu - (unsigned)(p & 0xFFFFFFFFull)) == a);
      match = in_s1 ? s1 : g2;
    } else {
      match = s1;
    }
  }

  float4 off = make_float4(0.f, 0.f, 0.f, 0.f);
  int lab = 0;  // background row = one_hot(0)
  if (match >= 0) {
    const float4 gb = __ldg(gt_xyxy + (size_t)b * Gmax + match);
    lab = __ldg(gt_label + (size_t)b * Gmax + match);
    const float4 d = __ldg(anc_cent + a);
    // datacoder.py:141-144 on the gathered GT corners
    const float gcx = __fmul_rn(__fadd_rn(gb.z, gb.x), 0.5f);
    const float gcy = __fmul_rn(__fadd_rn(gb.w, gb.y), 0.5f);
    const float gw = __fadd_rn(__fsub_rn(gb.z, gb.x), 1.0f);
    const float gh = __fadd_rn(__fsub_rn(gb.w, gb.y), 1.0f);
    // datacoder.py:266-269: two successive divisions; +1.0 inside the log
    off.x = __fdiv_rn(__fdiv_rn(__fsub_rn(gcx, d.x), d.z), sx);
    off.y = __fdiv_rn(__fdiv_rn(__fsub_rn(gcy, d.y), d.w), sy);
    off.z = __fdiv_rn(logf(__fadd_rn(__fdiv_rn(gw, d.z), 1.0f)), sw);
    off.w = __fdiv_rn(logf(__fadd_rn(__fdiv_rn(gh, d.w), 1.0f)), sh);
  }
  if (a < A) {
    st_stream_f4(out_offsets + (size_t)b * A + a, off);
    if (out_match) out_match[(size_t)b * A + a] = match;
  }

  const int n_here = min(kEmitThreads, A - a_base);
  if (C == 4) {
    if (a < A) {
      const float4 oh = make_float4(lab == 0 ? 1.f : 0.f, lab == 1 ? 1.f : 0.f, lab == 2 ? 1.f : 0.f,
                                    lab == 3 ? 1.f : 0.f);
      st_stream_f4(reinterpret_cast<float4*>(out_labels) + (size_t)b * A + a, oh);
    }
  } else {
    // generic class count: flat coalesced walk over the tile's (n_here * C) floats
    s_lab[tid] = lab;
    __syncthreads();
    float* dst = out_labels + ((size_t)b * A + a_base) * C;
    const int total = n_here * C;
    for (int e = tid; e < total; e += kEmitThreads) {
      const int al = e / C;
      const int c = e - al * C;
      dst[e] = (s_lab[al] == c) ? 1.0f : 0.0f;
    }
  }
}

}  // namespace

extern "C" size_t ssdx_encode_workspace_bytes(int32_t num_anchors, int32_t batch, int32_t gmax) {
  if (num_anchors <= 0 || batch <= 0 || gmax < 0) return 0;
  const size_t best = ssdx_align_up((size_t)batch * (size_t)(gmax > 0 ? gmax : 1) * sizeof(unsigned long long), 256);
  const size_t g2 = ssdx_align_up((size_t)batch * (size_t)num_anchors * sizeof(int), 256);
  return best + g2;
}

extern "C" int ssdx_encode(const ssdx_anchors* anchors, const float* gt_xyxy, const int32_t* gt_label,
                           const int32_t* gt_count, int32_t batch, int32_t gmax, int32_t num_classes,
                           float iou_threshold, const float* std_host, float* out_labels, float* out_offsets,
                           int32_t* out_match, void* workspace, size_t workspace_bytes, ssdx_stream stream) {
  SSDX_CHECK_ARG(anchors != nullptr, "ssdx_encode: anchors is NULL");
  SSDX_CHECK_ARG(batch >= 0 && gmax >= 0, "ssdx_encode: batch and gmax must be >= 0");
  SSDX_CHECK_ARG(num_classes >= 1, "ssdx_encode: num_classes must be >= 1");
  SSDX_CHECK_ARG(std_host != nullptr && out_labels != nullptr && out_offsets != nullptr && gt_count != nullptr,
                 "ssdx_encode: NULL pointer argument");
  SSDX_CHECK_ARG(gmax == 0 || (gt_xyxy != nullptr && gt_label != nullptr), "ssdx_encode: GT arrays are NULL");
  SSDX_CHECK_ARG(((uintptr_t)gt_xyxy & 15) == 0 && ((uintptr_t)out_offsets & 15) == 0 &&
                     ((uintptr_t)out_labels & 15) == 0,
                 "ssdx_encode: gt_xyxy/out_labels/out_offsets must be 16-byte aligned");
  SSDX_CHECK_ARG(batch <= 65535, "ssdx_encode: batch must be <= 65535 per call (got %d)", batch);
  if (batch == 0) return SSDX_OK;
  const int A = anchors->A;
  const size_t need = ssdx_encode_workspace_bytes(A, batch, gmax);
  if (workspace == nullptr || workspace_bytes < need) {
    ssdx_set_error("ssdx_encode: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return SSDX_ERR_WORKSPACE;
  }
  SSDX_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "ssdx_encode: workspace must be 256-byte aligned");
  SSDX_CUDA(cudaSetDevice(anchors->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int gm = gmax > 0 ? gmax : 1;
  const size_t best_bytes = ssdx_align_up((size_t)batch * gm * sizeof(unsigned long long), 256);
  unsigned long long* gt_best = (unsigned long long*)workspace;
  int* anchor_g2 = (int*)((char*)workspace + best_bytes);
  SSDX_CUDA(cudaMemsetAsync(gt_best, 0, (size_t)batch * gm * sizeof(unsigned long long), s));

  dim3 g1((A + kMatchThreads * kApt - 1) / (kMatchThreads * kApt), batch);
  ssdx_encode_match_kernel<<<g1, kMatchThreads, 0, s>>>(anchors->corners, anchors->area, (const float4*)gt_xyxy,
                                                        gt_count, A, gm, iou_threshold, gt_best, anchor_g2);
  SSDX_LAUNCH_CHECK("ssdx_encode_match_kernel");
  dim3 g2((A + kEmitThreads - 1) / kEmitThreads, batch);
  ssdx_encode_emit_kernel<<<g2, kEmitThreads, 0, s>>>(anchors->enc_cent, (const float4*)gt_xyxy, gt_label, gt_count,
                                                      A, gm, num_classes, std_host[0], std_host[1], std_host[2],
                                                      std_host[3], gt_best, anchor_g2, out_labels,
                                                      (float4*)out_offsets, out_match);
  SSDX_LAUNCH_CHECK("ssdx_encode_emit_kernel");
  return SSDX_OK;
}
