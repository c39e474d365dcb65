// Decode of standardised centroid offsets (reference: layers.py:57-79, datacoder.py:367-388, :408-432).
//
// Pure streaming: 16 B in + 16 B out per anchor per image => HBM-bound (DESIGN.md "decode").
// Each thread owns UNROLL anchors (centroids held in registers) and walks the batch with a
// grid-stride over images, so the anchor table is read once per CTA, the offsets are read with
// 128-bit streaming loads and the boxes are written with 128-bit streaming stores.
// fp32 association follows the reference: ((t*s)*d)+c, (exp(t*s)-1)*d, c -/+ (s-1)/2.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kUnroll = 2;

template <int MODE>
__device__ __forceinline__ float4 decode_one(const float4 t, const float4 d, const float sx, const float sy,
                                             const float sw, const float sh) {
  if (MODE == SSDX_DECODE_YXYX) return ssdx_decode_yxyx(t, d, sx, sy, sw, sh);  // shared with the NMS kernels
  // explicit _rn intrinsics: never contracted into FMA
  float cx = __fadd_rn(__fmul_rn(__fmul_rn(t.x, sx), d.z), d.x);
  float cy = __fadd_rn(__fmul_rn(__fmul_rn(t.y, sy), d.w), d.y);
  float w = __fmul_rn(__fsub_rn(expf(__fmul_rn(t.z, sw)), 1.0f), d.z);
  float h = __fmul_rn(__fsub_rn(expf(__fmul_rn(t.w, sh)), 1.0f), d.w);
  if (MODE != SSDX_DECODE_YXYX) {
    // datacoder.py:377-382: rows with sum(|offsets|) == 0 are multiplied by 0
    float s = __fadd_rn(__fadd_rn(__fadd_rn(fabsf(t.x), fabsf(t.y)), fabsf(t.z)), fabsf(t.w));
    float nb = (s > 0.0f) ? 1.0f : 0.0f;
    cx = __fmul_rn(cx, nb);
    cy = __fmul_rn(cy, nb);
    w = __fmul_rn(w, nb);
    h = __fmul_rn(h, nb);
    if (MODE == SSDX_DECODE_CENTROIDS_MASKED) return make_float4(cx, cy, w, h);
  }
  const float hw = __fmul_rn(__fsub_rn(w, 1.0f), 0.5f);  // (w-1)/2 exactly
  const float hh = __fmul_rn(__fsub_rn(h, 1.0f), 0.5f);
  float x0 = __fsub_rn(cx, hw), y0 = __fsub_rn(cy, hh), x1 = __fadd_rn(cx, hw), y1 = __fadd_rn(cy, hh);
  if (MODE == SSDX_DECODE_YXYX) return make_float4(y0, x0, y1, x1);  // layers.py:79
  // datacoder.py:420-426: masked again by sum(|centroids|) > 0
  float s2 = __fadd_rn(__fadd_rn(__fadd_rn(fabsf(cx), fabsf(cy)), fabsf(w)), fabsf(h));
  float nb2 = (s2 > 0.0f) ? 1.0f : 0.0f;
  return make_float4(__fmul_rn(x0, nb2), __fmul_rn(y0, nb2), __fmul_rn(x1, nb2), __fmul_rn(y1, nb2));
}

template <int MODE>
__global__ void __launch_bounds__(kThreads) ssdx_decode_kernel(const float4* __restrict__ anchors_cent,
                                                                const float4* __restrict__ offsets,
                                                                float4* __restrict__ out, const int A, const int B,
                                                                const float sx, const float sy, const float sw,
                                                                const float sh) {
  const int a0 = blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
  float4 d[kUnroll];
  bool ok[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const int a = a0 + u * kThreads;
    ok[u] = a < A;
    d[u] = ok[u] ? __ldg(anchors_cent + a) : make_float4(0.f, 0.f, 1.f, 1.f);
  }
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    const size_t row = (size_t)b * (size_t)A;
    float4 t[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (ok[u]) t[u] = ld_stream_f4(offsets + row + a0 + u * kThreads);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (ok[u]) st_stream_f4(out + row + a0 + u * kThreads, decode_one<MODE>(t[u], d[u], sx, sy, sw, sh));
  }
}

}  // namespace

extern "C" int ssdx_decode(const ssdx_anchors* anchors, const float* offsets, int32_t batch, const float* std_host,
                           int mode, float* out, ssdx_stream stream) {
  SSDX_CHECK_ARG(anchors != nullptr, "ssdx_decode: anchors is NULL");
  SSDX_CHECK_ARG(offsets != nullptr && out != nullptr && std_host != nullptr, "ssdx_decode: NULL pointer argument");
  SSDX_CHECK_ARG(batch >= 0, "ssdx_decode: batch must be >= 0");
  SSDX_CHECK_ARG(mode >= 0 && mode <= 2, "ssdx_decode: unknown mode %d", mode);
  SSDX_CHECK_ARG(((uintptr_t)offsets & 15) == 0 && ((uintptr_t)out & 15) == 0,
                 "ssdx_decode: offsets/out must be 16-byte aligned");
  if (batch == 0) return SSDX_OK;
  SSDX_DEVICE(anchors->device);
  const int A = anchors->A;
  const int gx = (A + kThreads * kUnroll - 1) / (kThreads * kUnroll);
  // ~16 resident CTAs per SM; every CTA walks >= 1 image
  const int target = ssdx_sm_count(anchors->device) * 16;
  int gy = (target + gx - 1) / gx;
  if (gy > batch) gy = batch;
  if (gy < 1) gy = 1;
  if (gy > 65535) gy = 65535;
  dim3 grid(gx, gy), block(kThreads);
  cudaStream_t s = (cudaStream_t)stream;
  const float4* cent = anchors->dec_cent;
  const float4* in = (const float4*)offsets;
  float4* o = (float4*)out;
  switch (mode) {
    case SSDX_DECODE_YXYX:
      ssdx_decode_kernel<SSDX_DECODE_YXYX><<<grid, block, 0, s>>>(cent, in, o, A, batch, std_host[0], std_host[1],
                                                                   std_host[2], std_host[3]);
      break;
    case SSDX_DECODE_CENTROIDS_MASKED:
      ssdx_decode_kernel<SSDX_DECODE_CENTROIDS_MASKED><<<grid, block, 0, s>>>(cent, in, o, A, batch, std_host[0],
                                                                               std_host[1], std_host[2], std_host[3]);
      break;
    default:
      ssdx_decode_kernel<SSDX_DECODE_CORNERS_MASKED><<<grid, block, 0, s>>>(cent, in, o, A, batch, std_host[0],
                                                                             std_host[1], std_host[2], std_host[3]);
      break;
  }
  SSDX_LAUNCH_CHECK("ssdx_decode_kernel");
  return SSDX_OK;
}
