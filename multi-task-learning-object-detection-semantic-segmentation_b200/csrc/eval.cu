// Detection evaluation: label-masked IoU of predictions vs ground truth, TP flags and per-class counts
// (reference: evaluators.py:6-63 _iou_boxes_pred_vs_true, :106-157 of average_precision_object_detection).
//
// One warp per image.  The GT boxes of the image are staged in shared memory (per-warp slice), every lane
// owns predictions k = lane, lane+32, ...  IoU uses the reference's inclusive-pixel convention with the
// 1e-7f epsilon in the denominator: inter / (((area_p + area_t) - inter) + 1e-7f), multiplied by the
// label-equality mask; TP iff label != background and max_g >= thr (fp32 compare; several predictions
// may be TP on the same GT box - the reference does no one-to-one assignment).
#include "common.cuh"

namespace {

constexpr int kEvalThreads = 128;  // 4 warps = 4 images per CTA
constexpr int kEvalGtTile = 64;    // GT boxes staged per warp per pass

__global__ void __launch_bounds__(kEvalThreads)
    ssdx_eval_tp_kernel(const int* __restrict__ pred_label, const float4* __restrict__ pred_xyxy,
                        const float4* __restrict__ gt_xyxy, const int* __restrict__ gt_label,
                        const int* __restrict__ gt_count, const int B, const int K, const int Gmax, const int C,
                        const int bg, const float thr, unsigned char* __restrict__ out_tp,
                        unsigned long long* __restrict__ counts) {
  __shared__ float4 s_box[kEvalThreads / 32][kEvalGtTile];
  __shared__ float s_area[kEvalThreads / 32][kEvalGtTile];
  __shared__ int s_lab[kEvalThreads / 32][kEvalGtTile];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * (kEvalThreads / 32) + warp;
  if (b >= B) return;  // whole warp exits together
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);

  // GT boxes per class (evaluators.py:121)
  for (int g = lane; g < G; g += 32) {
    const int l = gt_label[(size_t)b * Gmax + g];
    if (l >= 0 && l < C) atomicAdd(counts + (size_t)l * 3 + 2, 1ull);
  }

  for (int k0 = 0; k0 < K; k0 += 32) {
    const int k = k0 + lane;
    const bool have = k < K;
    const int pl = have ? pred_label[(size_t)b * K + k] : bg;
    const float4 pb = have ? __ldg(pred_xyxy + (size_t)b * K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float pa = __fmul_rn(__fadd_rn(__fsub_rn(pb.z, pb.x), 1.0f), __fadd_rn(__fsub_rn(pb.w, pb.y), 1.0f));
    // no GT: the reference scores a (P,1) zero matrix => max = 0
    float best = (G == 0) ? 0.0f : -INFINITY;
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
        const float4 tb = s_box[warp][g];
        const float iw = fmaxf(0.0f, __fadd_rn(__fsub_rn(fminf(pb.z, tb.z), fmaxf(pb.x, tb.x)), 1.0f));
        const float ih = fmaxf(0.0f, __fadd_rn(__fsub_rn(fminf(pb.w, tb.w), fmaxf(pb.y, tb.y)), 1.0f));
        const float inter = __fmul_rn(iw, ih);
        const float den = __fadd_rn(__fsub_rn(__fadd_rn(pa, s_area[warp][g]), inter), 1e-7f);
        float iou = __fdiv_rn(inter, den);
        iou = __fmul_rn(iou, (s_lab[warp][g] == pl) ? 1.0f : 0.0f);  // evaluators.py:61
        best = fmaxf(best, iou);
      }
    }
    const bool scored = have && pl != bg;  // evaluators.py:129-132
    const bool tp = scored && (best >= thr);
    if (have) out_tp[(size_t)b * K + k] = tp ? 1 : 0;
    // per-class TP / FP counts: one atomic per (warp, class) via match_any
    if (C <= 32) {
      const unsigned active = __ballot_sync(0xFFFFFFFFu, scored && pl >= 0 && pl < C);
      if (scored && pl >= 0 && pl < C) {
        const unsigned peers = __match_any_sync(active, pl);
        const unsigned tps = __ballot_sync(active, tp) & peers;
        if (lane == (__ffs(peers) - 1)) {
          const int ntp = __popc(tps), nall = __popc(peers);
          if (ntp) atomicAdd(counts + (size_t)pl * 3 + 0, (unsigned long long)ntp);
          if (nall - ntp) atomicAdd(counts + (size_t)pl * 3 + 1, (unsigned long long)(nall - ntp));
        }
      }
    } else if (scored && pl >= 0 && pl < C) {
      atomicAdd(counts + (size_t)pl * 3 + (tp ? 0 : 1), 1ull);
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
  SSDX_CUDA(cudaSetDevice(device));
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
  SSDX_CUDA(cudaSetDevice(device));
  const int per = kEvalThreads / 32;
  ssdx_eval_tp_kernel<<<(batch + per - 1) / per, kEvalThreads, 0, (cudaStream_t)stream>>>(
      pred_label, (const float4*)pred_xyxy, (const float4*)gt_xyxy, gt_label, gt_count, batch, k,
      gmax > 0 ? gmax : 1, num_classes, background_label, iou_threshold, out_tp, (unsigned long long*)counts);
  SSDX_LAUNCH_CHECK("ssdx_eval_tp_kernel");
  return SSDX_OK;
}
