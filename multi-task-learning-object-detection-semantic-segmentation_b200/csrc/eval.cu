// Detection evaluation: label-masked IoU of predictions vs ground truth, TP flags and per-class counts
// (reference: evaluators.py:6-63 _iou_boxes_pred_vs_true, :106-157 of average_precision_object_detection).
//
// One warp per image (persistent CTAs, warps stride over the batch).  Per-class counts are accumulated in shared memory
// per CTA and flushed with one global atomic per non-zero counter (at batch 4096 the direct version spent its time in
// ~240 k atomics on 12 addresses: 123 us -> see DESIGN.md 5.4).  The GT boxes of the image are staged in shared memory (per-warp slice), every lane
// owns predictions k = lane, lane+32, ...  IoU uses the reference's inclusive-pixel convention with the
// 1e-7f epsilon in the denominator: inter / (((area_p + area_t) - inter) + 1e-7f), multiplied by the
// label-equality mask; TP iff label != background and max_g >= thr (fp32 compare; several predictions
// may be TP on the same GT box - the reference does no one-to-one assignment).
#include "common.cuh"

namespace {

constexpr int kEvalThreads = 512;  // 16 warps per CTA, few CTAs: the per-CTA counter flush is what hits global atomics
constexpr int kEvalGtTile = 64;    // GT boxes staged per warp per pass

__global__ void __launch_bounds__(kEvalThreads)
    ssdx_eval_tp_kernel(const int* __restrict__ pred_label, const float4* __restrict__ pred_xyxy,
                        const float4* __restrict__ gt_xyxy, const int* __restrict__ gt_label,
                        const int* __restrict__ gt_count, const int B, const int K, const int Gmax, const int C,
                        const int bg, const float thr, unsigned char* __restrict__ out_tp,
                        unsigned long long* __restrict__ g_counts, const int smem_counts) {
  __shared__ float4 s_box[kEvalThreads / 32][kEvalGtTile];
  __shared__ float s_area[kEvalThreads / 32][kEvalGtTile];
  __shared__ int s_lab[kEvalThreads / 32][kEvalGtTile];
  extern __shared__ unsigned int s_cnt[];  // [C][3] when smem_counts
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (smem_counts) {
    for (int i = threadIdx.x; i < C * 3; i += kEvalThreads) s_cnt[i] = 0u;
    __syncthreads();
  }
  // counter sink: CTA-local (flushed at the end) or global
  auto count_add = [&](int cls, int which, unsigned n) {
    if (smem_counts) atomicAdd(&s_cnt[cls * 3 + which], n);
    else atomicAdd(g_counts + (size_t)cls * 3 + which, (unsigned long long)n);
  };
  // a warp task = (image, chunk of 32 predictions); an image with no predictions still has one task (GT counts)
  const int n_chunks = K > 0 ? (K + 31) / 32 : 1;
  const long long n_tasks = (long long)B * n_chunks;
  for (long long task = (long long)blockIdx.x * (kEvalThreads / 32) + warp; task < n_tasks;
       task += (long long)gridDim.x * (kEvalThreads / 32)) {
  const int b = (int)(task / n_chunks);
  const int k_first = (int)(task - (long long)b * n_chunks) * 32;
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);

  // GT boxes per class (evaluators.py:121), once per image
  for (int g0 = 0; g0 < G && k_first == 0; g0 += 32) {
    const int g = g0 + lane;
    const int l = (g < G) ? gt_label[(size_t)b * Gmax + g] : -1;
    const bool ok = l >= 0 && l < C;
    const unsigned act = __ballot_sync(0xFFFFFFFFu, ok);
    if (ok) {
      const unsigned peers = __match_any_sync(act, l);
      if (lane == (__ffs(peers) - 1)) count_add(l, 2, (unsigned)__popc(peers));
    }
  }

  for (int k0 = k_first; k0 < min(K, k_first + 32); k0 += 32) {
    const int k = k0 + lane;
    const bool have = k < K;
    const int pl = have ? pred_label[(size_t)b * K + k] : bg;
    const float4 pb = have ? __ldg(pred_xyxy + (size_t)b * K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float pa = __fmul_rn(__fadd_rn(__fsub_rn(pb.z, pb.x), 1.0f), __fadd_rn(__fsub_rn(pb.w, pb.y), 1.0f));
    // TP <=> max_g (iou * [labels equal]) >= thr (evaluators.py:149-153).  Only that decision is needed, so the
    // division is kept for the rare pair whose inter / den lies within 1e-6 of the threshold; everywhere else
    // inter vs thr * den decides with a margin far above the rounding of either side.  A GT box of another class
    // contributes iou * 0 = 0, and an image without GT scores a zero matrix (:27-28): both "hit" iff 0 >= thr.
    const bool zero_hits = 0.0f >= thr;
    bool hit = (G == 0) ? zero_hits : false;
    for (int g0 = 0; g0 < G; g0 += kEvalGtTile) {
      const int gn = min(kEvalGtTile, G - g0);
      __syncwarp();
      for (int g = lane; g < gn; g += 32) {
        const float4 tb = __ldg(gt_xyxy + (size_t)b * Gmax + g0 + g);
        s_box[warp][g] = tb;
        s_area[warp][g] = __fmul_rn(__fadd_rn(__fsub_rn(tb.z, tb.x), 1.0f), __fadd_rn(__fsub_rn(tb.w, tb.y), 1.0f));
        s_lab[warp][g] = gt_label[(size_t)b * Gmax + g0 + g];
      }
      __syncwarp();
      for (int g = 0; g < gn; ++g) {
        if (s_lab[warp][g] != pl) {
          hit = hit || zero_hits;
          continue;
        }
        const float4 tb = s_box[warp][g];
        const float iw = fmaxf(0.0f, __fadd_rn(__fsub_rn(fminf(pb.z, tb.z), fmaxf(pb.x, tb.x)), 1.0f));
        const float ih = fmaxf(0.0f, __fadd_rn(__fsub_rn(fminf(pb.w, tb.w), fmaxf(pb.y, tb.y)), 1.0f));
        const float inter = __fmul_rn(iw, ih);
        const float den = __fadd_rn(__fsub_rn(__fadd_rn(pa, s_area[warp][g]), inter), 1e-7f);
        const float rhs = __fmul_rn(thr, den);
        bool ge;
        if (!(den > 0.0f)) ge = __fdiv_rn(inter, den) >= thr;  // degenerate prediction (negative area): no shortcut
        else if (inter < __fmul_rn(rhs, rhs >= 0.0f ? 0.999999f : 1.000001f)) ge = false;
        else if (inter > __fmul_rn(rhs, rhs >= 0.0f ? 1.000001f : 0.999999f)) ge = true;
        else ge = __fdiv_rn(inter, den) >= thr;  // evaluators.py:58, exact
        hit = hit || ge;
      }
    }
    const bool scored = have && pl != bg;  // evaluators.py:129-132
    const bool tp = scored && hit;
    if (have) out_tp[(size_t)b * K + k] = tp ? 1 : 0;
    // per-class TP / FP counts: one atomic per (warp, class) via match_any
    if (C <= 32) {
      const unsigned active = __ballot_sync(0xFFFFFFFFu, scored && pl >= 0 && pl < C);
      if (scored && pl >= 0 && pl < C) {
        const unsigned peers = __match_any_sync(active, pl);
        const unsigned tps = __ballot_sync(active, tp) & peers;
        if (lane == (__ffs(peers) - 1)) {
          const int ntp = __popc(tps), nall = __popc(peers);
          if (ntp) count_add(pl, 0, (unsigned)ntp);
          if (nall - ntp) count_add(pl, 1, (unsigned)(nall - ntp));
        }
      }
    } else if (scored && pl >= 0 && pl < C) {
      count_add(pl, tp ? 0 : 1, 1u);
    }
  }
  }  // images of this warp
  if (smem_counts) {
    __syncthreads();
    for (int i = threadIdx.x; i < C * 3; i += kEvalThreads) {
      const unsigned v = s_cnt[i];
      if (v) atomicAdd(g_counts + i, (unsigned long long)v);
    }
  }
}

// (P,G) IoU matrix of evaluators.py:6-63; one thread per element, G fastest (coalesced stores)
__global__ void ssdx_iou_matrix_kernel(const int* __restrict__ pred_label, const float4* __restrict__ pred_xyxy,
                                       const int* __restrict__ gt_label, const float4* __restrict__ gt_xyxy,
                                       const int P, const int G, float* __restrict__ out) {
  const size_t n = (size_t)P * G;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(e / G), g = (int)(e - (size_t)p * G);
    const float4 pb = __ldg(pred_xyxy + p), tb = __ldg(gt_xyxy + g);
    const float pa = __fmul_rn(__fadd_rn(__fsub_rn(pb.z, pb.x), 1.0f), __fadd_rn(__fsub_rn(pb.w, pb.y), 1.0f));
    const float ta = __fmul_rn(__fadd_rn(__fsub_rn(tb.z, tb.x), 1.0f), __fadd_rn(__fsub_rn(tb.w, tb.y), 1.0f));
    const float iw = fmaxf(0.0f, __fadd_rn(__fsub_rn(fminf(pb.z, tb.z), fmaxf(pb.x, tb.x)), 1.0f));
    const float ih = fmaxf(0.0f, __fadd_rn(__fsub_rn(fminf(pb.w, tb.w), fmaxf(pb.y, tb.y)), 1.0f));
    const float inter = __fmul_rn(iw, ih);
    const float iou = __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(pa, ta), inter), 1e-7f));
    out[e] = __fmul_rn(iou, (pred_label[p] == gt_label[g]) ? 1.0f : 0.0f);
  }
}

}  // namespace

extern "C" int ssdx_iou_pred_vs_true(const int32_t* pred_label, const float* pred_xyxy, const int32_t* gt_label,
                                     const float* gt_xyxy, int32_t num_pred, int32_t num_gt, float* out, int device,
                                     ssdx_stream stream) {
  SSDX_CHECK_ARG(num_pred >= 0 && num_gt >= 0, "ssdx_iou_pred_vs_true: negative size");
  if (num_pred == 0 || num_gt == 0) return SSDX_OK;
  SSDX_CHECK_ARG(pred_label && pred_xyxy && gt_label && gt_xyxy && out, "ssdx_iou_pred_vs_true: NULL pointer");
  SSDX_CHECK_ARG(((uintptr_t)pred_xyxy & 15) == 0 && ((uintptr_t)gt_xyxy & 15) == 0,
                 "ssdx_iou_pred_vs_true: box arrays must be 16-byte aligned");
  SSDX_DEVICE(device);
  const size_t n = (size_t)num_pred * num_gt;
  int blocks = (int)((n + 255) / 256);
  const int cap = ssdx_sm_count(device) * 8;
  if (blocks > cap) blocks = cap;
  ssdx_iou_matrix_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred_label, (const float4*)pred_xyxy, gt_label,
                                                                   (const float4*)gt_xyxy, num_pred, num_gt, out);
  SSDX_LAUNCH_CHECK("ssdx_iou_matrix_kernel");
  return SSDX_OK;
}

extern "C" int ssdx_eval_tp(const int32_t* pred_label, const float* pred_xyxy, const float* gt_xyxy,
                            const int32_t* gt_label, const int32_t* gt_count, int32_t batch, int32_t k, int32_t gmax,
                            int32_t num_classes, int32_t background_label, float iou_threshold, uint8_t* out_tp,
                            int64_t* counts, int device, ssdx_stream stream) {
  SSDX_CHECK_ARG(batch >= 0 && k >= 0 && gmax >= 0 && num_classes >= 1, "ssdx_eval_tp: bad shape");
  SSDX_CHECK_ARG(out_tp != nullptr && counts != nullptr && gt_count != nullptr, "ssdx_eval_tp: NULL output/count");
  SSDX_CHECK_ARG(k == 0 || (pred_label != nullptr && pred_xyxy != nullptr), "ssdx_eval_tp: predictions are NULL");
  SSDX_CHECK_ARG(gmax == 0 || (gt_xyxy != nullptr && gt_label != nullptr), "ssdx_eval_tp: GT arrays are NULL");
  SSDX_CHECK_ARG(((uintptr_t)pred_xyxy & 15) == 0 && ((uintptr_t)gt_xyxy & 15) == 0,
                 "ssdx_eval_tp: box arrays must be 16-byte aligned");
  if (batch == 0) return SSDX_OK;
  SSDX_DEVICE(device);
  const int per = kEvalThreads / 32;
  const int smem_counts = num_classes <= 2048;  // 24 KB of CTA-local counters at most
  const long long tasks = (long long)batch * (k > 0 ? (k + 31) / 32 : 1);
  int grid = (int)((tasks + per - 1) / per > (1ll << 30) ? (1ll << 30) : (tasks + per - 1) / per);
  const int cap = ssdx_sm_count(device) * 2;  // persistent: 2 CTAs of 16 warps per SM
  if (smem_counts && grid > cap) grid = cap;
  // (CTA-local 32-bit counters cannot overflow: a CTA sees at most batch * (k + gmax) / grid events < 2^32 here)
  SSDX_CHECK_ARG(!smem_counts || ((tasks + grid - 1) / grid) * per * (32ll + gmax) < (1ll << 32),
                 "ssdx_eval_tp: batch too large for one call");
  ssdx_eval_tp_kernel<<<grid, kEvalThreads, smem_counts ? (size_t)num_classes * 3 * sizeof(unsigned) : 0,
                        (cudaStream_t)stream>>>(
      pred_label, (const float4*)pred_xyxy, (const float4*)gt_xyxy, gt_label, gt_count, batch, k,
      gmax > 0 ? gmax : 1, num_classes, background_label, iou_threshold, out_tp, (unsigned long long*)counts,
      smem_counts);
  SSDX_LAUNCH_CHECK("ssdx_eval_tp_kernel");
  return SSDX_OK;
}
