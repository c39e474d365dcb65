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
// The A x G IoU matrix is never materialised.  The pair loop (match_tile) is shared by two drivers:
//   fused  : one CTA per image (A and G small enough for shared memory: the reference configuration).
//            Phase 1 walks the anchors in warp tiles of 32*kApt, keeps the per-anchor arg-max in
//            registers -> int16 table in shared memory, and the per-GT arg-max per warp (REDUX on the IoU
//            bit pattern, ties -> REDUX.MIN of the anchor index).  Phase 2 resolves the closed form and
//            streams labels + offsets out with 128-bit coalesced stores.  No workspace traffic at all.
//   tiled  : grid (anchor tiles, images) + a second emit kernel, for large A (dense stress: 100k anchors);
//            the per-GT arg-max crosses CTAs through one 64-bit atomicMax per (CTA, GT).
// Work reduction that keeps bit-exactness: (i) a warp tile whose bounding box misses the GT box skips its
// 32*kApt pairs with one warp-uniform test (fp32 rounding is monotone, so the tile test is conservative);
// (ii) pairs with empty intersection skip the division: their IoU is exactly +0 and can never win a strict
// '>' comparison against a running best that starts at (0, index 0).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kApt = 4;             // anchors per thread per tile
constexpr int kFusedThreads = 512;  // one CTA per image
constexpr int kTiledThreads = 256;  // grid (tiles, images)
constexpr int kGtTile = 256;        // GT boxes staged per shared-memory tile (tiled driver)
constexpr int kEmitThreads = 256;
constexpr int kFusedMaxG = 256;

typedef unsigned long long u64;

// packed (iou bits, inverted anchor index): larger = better IoU, ties -> lower anchor index
__device__ __forceinline__ u64 pack_best(unsigned iou_bits, int a) {
  return ((u64)iou_bits << 32) | (u64)(0xFFFFFFFFu - (unsigned)a);
}
__device__ __forceinline__ int best_anchor(u64 p) { return (int)(0xFFFFFFFFu - (unsigned)(p & 0xFFFFFFFFull)); }

__device__ __forceinline__ float gt_area(const float4 b) {
  // datacoder.py:206  (xmax-xmin+1)*(ymax-ymin+1)
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
}

struct AnchorTile {
  float4 box[kApt];
  float area[kApt];
  int idx[kApt];
  float best_v[kApt];
  int best_g[kApt];
  float4 bbox;  // warp-uniform bounding box of the 32*kApt anchors
};

// lane l of warp tile starting at `base` owns anchors base + u*32 + l (ascending in u)
__device__ __forceinline__ void load_tile(AnchorTile& t, const float4* __restrict__ corners,
                                          const float* __restrict__ area, int base, int lane, int A) {
  const float inf = __int_as_float(0x7f800000);
  float4 bb = make_float4(inf, inf, -inf, -inf);
#pragma unroll
  for (int u = 0; u < kApt; ++u) {
    const int a = base + u * 32 + lane;
    t.idx[u] = a;
    const bool ok = a < A;
    // out-of-range slot: empty box, inter == 0 against everything, neutral for the bounding box
    t.box[u] = ok ? __ldg(corners + a) : make_float4(inf, inf, -inf, -inf);
    t.area[u] = ok ? __ldg(area + a) : 1.0f;
    t.best_v[u] = 0.0f;  // np.argmax over an all-zero row -> index 0
    t.best_g[u] = 0;
    bb.x = fminf(bb.x, t.box[u].x);
    bb.y = fminf(bb.y, t.box[u].y);
    bb.z = fmaxf(bb.z, t.box[u].z);
    bb.w = fmaxf(bb.w, t.box[u].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    bb.x = fminf(bb.x, __shfl_xor_sync(0xFFFFFFFFu, bb.x, o));
    bb.y = fminf(bb.y, __shfl_xor_sync(0xFFFFFFFFu, bb.y, o));
    bb.z = fmaxf(bb.z, __shfl_xor_sync(0xFFFFFFFFu, bb.z, o));
    bb.w = fmaxf(bb.w, __shfl_xor_sync(0xFFFFFFFFu, bb.w, o));
  }
  t.bbox = bb;
}

// All pairs (anchors of the tile) x (gn GT boxes in shared memory).  wrow[g] is this warp's running packed best.
// (shared-memory pointers: no __restrict__, see nms.cu)
__device__ __forceinline__ void match_tile(AnchorTile& t, const float4* s_box, const float* s_area, const int gn,
                                           const int g0, u64* wrow, const int lane) {
  for (int g = 0; g < gn; ++g) {
    const float4 gb = s_box[g];
    // warp-uniform cull: the same expression as the pair intersection, evaluated on the tile's bounding box
    const float tw = __fadd_rn(__fsub_rn(fminf(t.bbox.z, gb.z), fmaxf(t.bbox.x, gb.x)), 1.0f);
    const float th = __fadd_rn(__fsub_rn(fminf(t.bbox.w, gb.w), fmaxf(t.bbox.y, gb.y)), 1.0f);
    if (!(tw > 0.0f) || !(th > 0.0f)) continue;
    const float ga = s_area[g];
    float mv = 0.0f;
    int ma = 0x7FFFFFFF;
#pragma unroll
    for (int u = 0; u < kApt; ++u) {
      // datacoder.py:210-216
      const float iw = __fadd_rn(__fsub_rn(fminf(t.box[u].z, gb.z), fmaxf(t.box[u].x, gb.x)), 1.0f);
      const float ih = __fadd_rn(__fsub_rn(fminf(t.box[u].w, gb.w), fmaxf(t.box[u].y, gb.y)), 1.0f);
      const float inter = __fmul_rn(fmaxf(0.0f, iw), fmaxf(0.0f, ih));
      if (inter > 0.0f) {
        // datacoder.py:220  inter / ((area_a + area_g) - inter), IEEE division
        const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(t.area[u], ga), inter));
        if (iou > t.best_v[u]) {  // strict: ties keep the lower GT index
          t.best_v[u] = iou;
          t.best_g[u] = g0 + g;
        }
        if (iou > mv) {  // strict: ties keep the lower anchor index (u ascending)
          mv = iou;
          ma = t.idx[u];
        }
      }
    }
    // warp arg-max for this GT: IoU >= 0 so the bit pattern is monotone
    const unsigned mbits = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(mv));
    if (mbits != 0u) {  // warp-uniform
      // lowest anchor index among the lanes that tie on the maximum (every lane takes part in the REDUX)
      const int amin = __reduce_min_sync(0xFFFFFFFFu, (__float_as_uint(mv) == mbits) ? ma : 0x7FFFFFFF);
      if (lane == 0) {
        const u64 wb = pack_best(mbits, amin);
        if (wb > wrow[g]) wrow[g] = wb;
      }
    }
  }
}

// closed form of concat -> unique -> sequential scatter for one anchor
__device__ __forceinline__ int resolve_match(int a, int g2, int s1max, const u64* gbest) {
  if (g2 >= 0) {
    const u64 p = gbest[g2];
    const bool in_s1 = ((p >> 32) != 0ull) && (best_anchor(p) == a);
    return in_s1 ? s1max : g2;
  }
  return s1max;
}

// datacoder.py:251-269 for one matched pair
__device__ __forceinline__ float4 pair_offsets(const float4 gb, const float4 d, float sx, float sy, float sw,
                                               float sh) {
  const float gcx = __fmul_rn(__fadd_rn(gb.z, gb.x), 0.5f);
  const float gcy = __fmul_rn(__fadd_rn(gb.w, gb.y), 0.5f);
  const float gw = __fadd_rn(__fsub_rn(gb.z, gb.x), 1.0f);
  const float gh = __fadd_rn(__fsub_rn(gb.w, gb.y), 1.0f);
  float4 off;  // two successive divisions; +1.0 inside the log
  off.x = __fdiv_rn(__fdiv_rn(__fsub_rn(gcx, d.x), d.z), sx);
  off.y = __fdiv_rn(__fdiv_rn(__fsub_rn(gcy, d.y), d.w), sy);
  off.z = __fdiv_rn(logf(__fadd_rn(__fdiv_rn(gw, d.z), 1.0f)), sw);
  off.w = __fdiv_rn(logf(__fadd_rn(__fdiv_rn(gh, d.w), 1.0f)), sh);
  return off;
}

// ------------------------------------------------------------------------------------------ fused driver
struct FusedSmem {
  float4* box;    // [G]
  u64* wbest;     // [warps][G]
  u64* gbest;     // [G]
  float* area;    // [G]
  int* lab;       // [G]
  short* g2;      // [A]  per-anchor arg-max GT (or -1); later: matched label
  short* s1;      // [A]  max GT index whose best anchor is this anchor (or -1)
};

__host__ __device__ inline size_t fused_smem_bytes(int A, int G) {
  const size_t g = (size_t)(G > 0 ? G : 1);
  return g * 16 + g * (kFusedThreads / 32) * 8 + g * 8 + g * 4 + g * 4 + (size_t)((A + 7) / 8 * 8) * 2 * 2 + 16;
}

__global__ void __launch_bounds__(kFusedThreads, 2)
    ssdx_encode_fused_kernel(const float4* __restrict__ anc_corners, const float* __restrict__ anc_area,
                             const float4* __restrict__ anc_cent, const float4* __restrict__ gt_xyxy,
                             const int* __restrict__ gt_label, const int* __restrict__ gt_count, const int A,
                             const int Gmax, const int C, const float iou_thr, const float sx, const float sy,
                             const float sw, const float sh, float* __restrict__ out_labels,
                             float4* __restrict__ out_offsets, int* __restrict__ out_match) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kFusedThreads / 32;
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);
  const int Gs = Gmax > 0 ? Gmax : 1;
  FusedSmem sm;
  {
    unsigned char* p = smem_raw;
    sm.box = (float4*)p;  p += (size_t)Gs * 16;
    sm.wbest = (u64*)p;   p += (size_t)Gs * kWarps * 8;
    sm.gbest = (u64*)p;   p += (size_t)Gs * 8;
    sm.area = (float*)p;  p += (size_t)Gs * 4;
    sm.lab = (int*)p;     p += (size_t)Gs * 4;
    sm.g2 = (short*)p;    p += (size_t)((A + 7) / 8 * 8) * 2;
    sm.s1 = (short*)p;
  }
  for (int g = tid; g < G; g += kFusedThreads) {
    const float4 bx = __ldg(gt_xyxy + (size_t)b * Gmax + g);
    sm.box[g] = bx;
    sm.area[g] = gt_area(bx);
    sm.lab[g] = __ldg(gt_label + (size_t)b * Gmax + g);
  }
  for (int i = tid; i < G * kWarps; i += kFusedThreads) sm.wbest[i] = 0ull;
  for (int i = tid; i < (A + 1) / 2; i += kFusedThreads) reinterpret_cast<int*>(sm.s1)[i] = -1;  // two int16 -1
  __syncthreads();

  // ---- phase 1: all pairs ----
  if (G > 0) {
    u64* wrow = sm.wbest + (size_t)warp * G;
    for (int base = warp * (32 * kApt); base < A; base += kWarps * 32 * kApt) {
      AnchorTile t;
      load_tile(t, anc_corners, anc_area, base, lane, A);
      match_tile(t, sm.box, sm.area, G, 0, wrow, lane);
#pragma unroll
      for (int u = 0; u < kApt; ++u)
        if (t.idx[u] < A) sm.g2[t.idx[u]] = (short)((t.best_v[u] > iou_thr) ? t.best_g[u] : -1);
    }
  } else {
    for (int i = tid; i < (A + 1) / 2; i += kFusedThreads) reinterpret_cast<int*>(sm.g2)[i] = -1;
  }
  __syncthreads();
  for (int g = tid; g < G; g += kFusedThreads) {
    u64 m = 0ull;
    for (int w = 0; w < kWarps; ++w) {
      const u64 v = sm.wbest[(size_t)w * G + g];
      m = v > m ? v : m;
    }
    sm.gbest[g] = m;
  }
  __syncthreads();
  if (tid == 0)  // ascending g: the last writer is max(S1(a)); G is small on this path
    for (int g = 0; g < G; ++g) {
      const u64 p = sm.gbest[g];
      if ((p >> 32) != 0ull) sm.s1[best_anchor(p)] = (short)g;
    }
  __syncthreads();

  // ---- phase 2: resolve + emit (coalesced 128-bit streaming stores) ----
  float4* offs = out_offsets + (size_t)b * A;
  for (int a = tid; a < A; a += kFusedThreads) {
    const int match = resolve_match(a, sm.g2[a], sm.s1[a], sm.gbest);
    float4 off = make_float4(0.f, 0.f, 0.f, 0.f);
    int lab = 0;  // background row = one_hot(0)
    if (match >= 0) {
      lab = sm.lab[match];
      off = pair_offsets(sm.box[match], __ldg(anc_cent + a), sx, sy, sw, sh);
    }
    st_stream_f4(offs + a, off);
    if (out_match) out_match[(size_t)b * A + a] = match;
    if (C == 4) {
      const float4 oh = make_float4(lab == 0 ? 1.f : 0.f, lab == 1 ? 1.f : 0.f, lab == 2 ? 1.f : 0.f,
                                    lab == 3 ? 1.f : 0.f);
      st_stream_f4(reinterpret_cast<float4*>(out_labels) + (size_t)b * A + a, oh);
    } else {
      // own slot only; a label outside [0, C) can never equal a class index, so clamping to int16 is safe
      sm.g2[a] = (short)((lab >= 0 && lab < 32767) ? lab : 32767);
    }
  }
  if (C != 4) {
    __syncthreads();
    float* dst = out_labels + (size_t)b * A * C;
    const int total = A * C;  // flat coalesced walk
    for (int e = tid; e < total; e += kFusedThreads) {
      const int a = e / C;
      const int c = e - a * C;
      dst[e] = ((int)sm.g2[a] == c) ? 1.0f : 0.0f;
    }
  }
}

// ------------------------------------------------------------------------------------------ tiled driver
__global__ void __launch_bounds__(kTiledThreads)
    ssdx_encode_match_kernel(const float4* __restrict__ anc_corners, const float* __restrict__ anc_area,
                             const float4* __restrict__ gt_xyxy, const int* __restrict__ gt_count, const int A,
                             const int Gmax, const float iou_thr, u64* __restrict__ gt_best,
                             int* __restrict__ anchor_g2) {
  __shared__ float4 s_box[kGtTile];
  __shared__ float s_area[kGtTile];
  __shared__ u64 s_wbest[(kTiledThreads / 32)][kGtTile];  // warp-private running best per GT

  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);
  const int base = blockIdx.x * (kTiledThreads * kApt) + warp * (32 * kApt);

  AnchorTile t;
  load_tile(t, anc_corners, anc_area, base, lane, A);

  for (int g0 = 0; g0 < G; g0 += kGtTile) {
    const int gn = min(kGtTile, G - g0);
    __syncthreads();
    for (int g = tid; g < gn; g += kTiledThreads) {
      const float4 bx = __ldg(gt_xyxy + (size_t)b * Gmax + g0 + g);
      s_box[g] = bx;
      s_area[g] = gt_area(bx);
    }
    for (int i = tid; i < kGtTile * (kTiledThreads / 32); i += kTiledThreads) (&s_wbest[0][0])[i] = 0ull;
    __syncthreads();
    match_tile(t, s_box, s_area, gn, g0, s_wbest[warp], lane);
    __syncthreads();
    // one global atomic per (CTA, GT) that saw a positive IoU
    for (int g = tid; g < gn; g += kTiledThreads) {
      u64 m = 0ull;
#pragma unroll
      for (int w = 0; w < kTiledThreads / 32; ++w) {
        const u64 v = s_wbest[w][g];
        m = v > m ? v : m;
      }
      if (m != 0ull) atomicMax(gt_best + (size_t)b * Gmax + g0 + g, m);
    }
  }
#pragma unroll
  for (int u = 0; u < kApt; ++u)
    if (t.idx[u] < A)
      anchor_g2[(size_t)b * A + t.idx[u]] = (G > 0 && t.best_v[u] > iou_thr) ? t.best_g[u] : -1;
}

__global__ void __launch_bounds__(kEmitThreads)
    ssdx_encode_emit_kernel(const float4* __restrict__ anc_cent, const float4* __restrict__ gt_xyxy,
                            const int* __restrict__ gt_label, const int* __restrict__ gt_count, const int A,
                            const int Gmax, const int C, const float sx, const float sy, const float sw,
                            const float sh, const u64* __restrict__ gt_best, const int* __restrict__ anchor_g2,
                            float* __restrict__ out_labels, float4* __restrict__ out_offsets,
                            int* __restrict__ out_match) {
  __shared__ int s_s1max[kEmitThreads];
  __shared__ int s_lab[kEmitThreads];
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int a_base = blockIdx.x * kEmitThreads;
  const int a = a_base + tid;
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);
  const u64* best = gt_best + (size_t)b * Gmax;

  s_s1max[tid] = -1;
  __syncthreads();
  // S1: GT boxes whose best anchor (IoU > 0) lies in this tile
  for (int g = tid; g < G; g += kEmitThreads) {
    const u64 p = best[g];
    if ((p >> 32) != 0ull) {
      const int ba = best_anchor(p);
      if (ba >= a_base && ba < a_base + kEmitThreads) atomicMax(&s_s1max[ba - a_base], g);
    }
  }
  __syncthreads();

  int match = -1;
  if (a < A) match = resolve_match(a, anchor_g2[(size_t)b * A + a], s_s1max[tid], best);
  float4 off = make_float4(0.f, 0.f, 0.f, 0.f);
  int lab = 0;  // background row = one_hot(0)
  if (match >= 0) {
    lab = __ldg(gt_label + (size_t)b * Gmax + match);
    off = pair_offsets(__ldg(gt_xyxy + (size_t)b * Gmax + match), __ldg(anc_cent + a), sx, sy, sw, sh);
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

inline bool use_fused(int A, int gmax, int batch, int sms) {
  // one CTA per image needs enough images to fill the GPU, and A/G small enough for shared memory
  const bool fits = gmax <= kFusedMaxG && A <= 32768 && fused_smem_bytes(A, gmax) <= 100 * 1024;
  const char* force = getenv("SSDX_ENCODE_PATH");  // "fused" | "tiled": test hook, both paths are bit-identical
  if (force != nullptr && force[0] == 't') return false;
  if (force != nullptr && force[0] == 'f') return fits;
  return fits && batch >= sms;
}

}  // namespace

extern "C" size_t ssdx_encode_workspace_bytes(int32_t num_anchors, int32_t batch, int32_t gmax) {
  if (num_anchors <= 0 || batch <= 0 || gmax < 0) return 0;
  const size_t best = ssdx_align_up((size_t)batch * (size_t)(gmax > 0 ? gmax : 1) * sizeof(u64), 256);
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
  SSDX_CHECK_ARG(gmax <= 32767, "ssdx_encode: at most 32767 GT boxes per image (got %d)", gmax);
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

  if (use_fused(A, gmax, batch, ssdx_sm_count(anchors->device))) {
    const size_t smem = fused_smem_bytes(A, gmax);
    SSDX_CUDA(cudaFuncSetAttribute(ssdx_encode_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ssdx_encode_fused_kernel<<<batch, kFusedThreads, smem, s>>>(
        anchors->corners, anchors->area, anchors->enc_cent, (const float4*)gt_xyxy, gt_label, gt_count, A, gmax,
        num_classes, iou_threshold, std_host[0], std_host[1], std_host[2], std_host[3], out_labels,
        (float4*)out_offsets, out_match);
    SSDX_LAUNCH_CHECK("ssdx_encode_fused_kernel");
    return SSDX_OK;
  }

  const size_t best_bytes = ssdx_align_up((size_t)batch * gm * sizeof(u64), 256);
  u64* gt_best = (u64*)workspace;
  int* anchor_g2 = (int*)((char*)workspace + best_bytes);
  SSDX_CUDA(cudaMemsetAsync(gt_best, 0, (size_t)batch * gm * sizeof(u64), s));
  dim3 g1((A + kTiledThreads * kApt - 1) / (kTiledThreads * kApt), batch);
  ssdx_encode_match_kernel<<<g1, kTiledThreads, 0, s>>>(anchors->corners, anchors->area, (const float4*)gt_xyxy,
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
