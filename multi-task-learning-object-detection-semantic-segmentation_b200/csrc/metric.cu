// Training-time box IoU metric (reference: metrics.py:53-173 jaccard_iou_bounding_boxes), SURVEY 8(f) row 3: a fused
// consumer of the encoded targets.  Per image: the mean, over the anchors whose target row is not background
// (sum |y_true| > 0, metrics.py:138-139), of the IoU between the decoded prediction and the decoded target, with the
// reference's own arithmetic (max(0, .) clamp on the decoded width/height :108-109, +1 pixel convention, Keras
// epsilon 1e-7 in the denominator :164).  Background rows contribute exactly +0 to the sum, so they are only counted
// out, not evaluated.  One CTA per image, two float4 streams in (32 B per anchor), one float out: HBM-bound.
#include "common.cuh"

namespace {

constexpr int kMetricThreads = 256;
constexpr int kMetricBatch = 4;  // anchors in flight per thread (8 independent 128-bit loads)

struct Decoded {
  float x0, y0, x1, y1, w, h;
};
// metrics.py:96-121 for a non-background row (not_background == 1: the multiplications by it are identities)
__device__ __forceinline__ Decoded decode_clamped(const float4 t, const float4 d, const float sx, const float sy,
                                                  const float sw, const float sh) {
  const float cx = __fadd_rn(__fmul_rn(__fmul_rn(t.x, sx), d.z), d.x);
  const float cy = __fadd_rn(__fmul_rn(__fmul_rn(t.y, sy), d.w), d.y);
  const float w = fmaxf(0.0f, __fmul_rn(__fsub_rn(expf(__fmul_rn(t.z, sw)), 1.0f), d.z));
  const float h = fmaxf(0.0f, __fmul_rn(__fsub_rn(expf(__fmul_rn(t.w, sh)), 1.0f), d.w));
  const float hw = __fmul_rn(__fsub_rn(w, 1.0f), 0.5f), hh = __fmul_rn(__fsub_rn(h, 1.0f), 0.5f);
  return Decoded{__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh), w, h};
}

__global__ void __launch_bounds__(kMetricThreads)
    ssdx_metric_iou_boxes_kernel(const float4* __restrict__ cent, const float4* __restrict__ y_true,
                                 const float4* __restrict__ y_pred, const int A, const float sx, const float sy,
                                 const float sw, const float sh, float* __restrict__ out) {
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* yt = y_true + (size_t)b * A;
  const float4* yp = y_pred + (size_t)b * A;
  float sum = 0.0f;
  int npos = 0;
  for (int a0 = tid; a0 < A; a0 += kMetricThreads * kMetricBatch) {
    float4 t[kMetricBatch], p[kMetricBatch];
#pragma unroll
    for (int u = 0; u < kMetricBatch; ++u) {
      const int a = a0 + u * kMetricThreads;
      const bool ok = a < A;
      t[u] = ok ? ld_stream_f4(yt + a) : make_float4(0.f, 0.f, 0.f, 0.f);
      p[u] = ok ? ld_stream_f4(yp + a) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kMetricBatch; ++u) {
      const float s = __fadd_rn(__fadd_rn(__fadd_rn(fabsf(t[u].x), fabsf(t[u].y)), fabsf(t[u].z)), fabsf(t[u].w));
      if (s > 0.0f) {  // padding slots are all-zero rows
        const float4 d = __ldg(cent + a0 + u * kMetricThreads);
        const Decoded P = decode_clamped(p[u], d, sx, sy, sw, sh);
        const Decoded T = decode_clamped(t[u], d, sx, sy, sw, sh);
        const float wi = fmaxf(0.0f, __fadd_rn(__fsub_rn(fminf(T.x1, P.x1), fmaxf(T.x0, P.x0)), 1.0f));
        const float hi = fmaxf(0.0f, __fadd_rn(__fsub_rn(fminf(T.y1, P.y1), fmaxf(T.y0, P.y0)), 1.0f));
        const float ai = __fmul_rn(wi, hi);
        const float at = __fmul_rn(T.w, T.h), ap = __fmul_rn(P.w, P.h);
        sum = __fadd_rn(sum, __fdiv_rn(ai, __fadd_rn(__fsub_rn(__fadd_rn(ap, at), ai), 1e-7f)));
        ++npos;
      }
    }
  }
  // CTA reduction (fp32 tree; the reference's reduce_sum order is unspecified as well)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum = __fadd_rn(sum, __shfl_xor_sync(0xFFFFFFFFu, sum, o));
  npos = __reduce_add_sync(0xFFFFFFFFu, npos);
  __shared__ float s_sum[kMetricThreads / 32];
  __shared__ int s_n[kMetricThreads / 32];
  if (lane == 0) {
    s_sum[warp] = sum;
    s_n[warp] = npos;
  }
  __syncthreads();
  if (tid == 0) {
    float tot = 0.0f;
    int n = 0;
    for (int w = 0; w < kMetricThreads / 32; ++w) {
      tot = __fadd_rn(tot, s_sum[w]);
      n += s_n[w];
    }
    out[b] = __fdiv_rn(tot, (float)n);  // 0 / 0 = NaN for an image without positives, as the reference
  }
}

}  // namespace

extern "C" int ssdx_metric_iou_boxes(const ssdx_anchors* anchors, const float* y_true, const float* y_pred,
                                     int32_t batch, const float* std_host, float* out, ssdx_stream stream) {
  SSDX_CHECK_ARG(anchors != nullptr, "ssdx_metric_iou_boxes: anchors is NULL");
  SSDX_CHECK_ARG(y_true != nullptr && y_pred != nullptr && std_host != nullptr && out != nullptr,
                 "ssdx_metric_iou_boxes: NULL pointer argument");
  SSDX_CHECK_ARG(batch >= 0, "ssdx_metric_iou_boxes: batch must be >= 0");
  SSDX_CHECK_ARG(((uintptr_t)y_true & 15) == 0 && ((uintptr_t)y_pred & 15) == 0,
                 "ssdx_metric_iou_boxes: y_true / y_pred must be 16-byte aligned");
  if (batch == 0) return SSDX_OK;
  SSDX_DEVICE(anchors->device);
  ssdx_metric_iou_boxes_kernel<<<batch, kMetricThreads, 0, (cudaStream_t)stream>>>(
      anchors->dec_cent, (const float4*)y_true, (const float4*)y_pred, anchors->A, std_host[0], std_host[1],
      std_host[2], std_host[3], out);
  SSDX_LAUNCH_CHECK("ssdx_metric_iou_boxes_kernel");
  return SSDX_OK;
}
