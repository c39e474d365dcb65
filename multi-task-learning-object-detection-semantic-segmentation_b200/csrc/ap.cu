// Average precision on the device (reference: evaluators.py:159-187), SURVEY 8(f) row 4: the tail of
// average_precision_object_detection after the TP flags.  Per foreground class: predictions sorted by confidence
// (descending), cum = cumsum(tp) (fp32, exact: counts stay below 2^24), precision = cum / rank (fp64),
// recall = cum / n_gt (fp32, as NumPy computes float32 / int), AP = trapezoid(precision over recall) in fp64 with no
// (0, .) sentinel; 0 when the class has no ground truth or no prediction.
// The sort is one cub::DeviceRadixSort over composite keys (class << 32 | ~confidence key): library code, like a
// cuBLAS call; the scan + trapezoid kernel is ours.  Ties in confidence are ordered arbitrarily by NumPy's quicksort
// in the reference; here by the radix sort's stability (input order).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace {

typedef unsigned long long u64;
constexpr int kApThreads = 512;

__device__ __forceinline__ unsigned conf_key_desc(float s) {  // smaller key = larger confidence
  const unsigned u = __float_as_uint(s);
  const unsigned asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~asc;
}

__global__ void ssdx_ap_keys_kernel(const unsigned char* __restrict__ tp, const float* __restrict__ conf,
                                    const int* __restrict__ label, const long long n, const int C, const int bg,
                                    u64* __restrict__ keys, unsigned* __restrict__ vals, int* __restrict__ class_count) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int l = label[i];
    const bool scored = l != bg && l >= 0 && l < C;
    const unsigned cls = scored ? (unsigned)l : (unsigned)C;  // unscored predictions sort behind every class
    keys[i] = ((u64)cls << 32) | (u64)conf_key_desc(conf[i]);
    vals[i] = tp[i] ? 1u : 0u;
    if (scored) atomicAdd(class_count + l, 1);
  }
}

// one CTA per class; segment [off, off + m) of the sorted arrays
__global__ void __launch_bounds__(kApThreads)
    ssdx_ap_scan_kernel(const unsigned* __restrict__ vals, const int* __restrict__ class_count,
                        const long long* __restrict__ counts /*C,3*/, const int C, const int bg,
                        double* __restrict__ out_ap) {
  const int c = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int s_warp[kApThreads / 32];
  __shared__ double s_red[kApThreads / 32];
  __shared__ int s_carry;
  long long off = 0;
  for (int k = 0; k < c; ++k) off += class_count[k];
  const int m = class_count[c];
  const long long n_gt = counts[(size_t)c * 3 + 2];
  if (c == bg || m == 0 || n_gt == 0) {  // evaluators.py:167-168
    if (tid == 0) out_ap[c] = 0.0;
    return;
  }
  const float ngt_f = (float)n_gt;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  double acc = 0.0;
  constexpr int kItems = 8;  // consecutive elements per thread per chunk
  for (int base = 0; base < m; base += kApThreads * kItems) {
    const int i0 = base + tid * kItems;
    int v[kItems + 1];  // v[kItems] = first element of the next thread (for the last trapezoid of this thread)
#pragma unroll
    for (int k = 0; k <= kItems; ++k) v[k] = (i0 + k < m) ? (int)vals[off + i0 + k] : 0;
    int tsum = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) tsum += v[k];
    // exclusive block scan of the per-thread sums
    int x = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    int wsum = 0;
    for (int w = 0; w < warp; ++w) wsum += s_warp[w];
    int cum = s_carry + wsum + x - tsum;  // cumsum(tp) before this thread's first element
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      const int i = i0 + k;
      const int cum_i = cum + v[k];  // cumsum(tp)[i]
      if (i + 1 < m) {               // trapezoid between points i and i + 1
        const int cum_n = cum_i + v[k + 1];
        const float r0 = __fdiv_rn((float)cum_i, ngt_f), r1 = __fdiv_rn((float)cum_n, ngt_f);
        const double p0 = (double)(float)cum_i / (double)(i + 1), p1 = (double)(float)cum_n / (double)(i + 2);
        acc += (double)__fsub_rn(r1, r0) * (p1 + p0) / 2.0;
      }
      cum = cum_i;
    }
    __syncthreads();
    if (tid == kApThreads - 1) s_carry = cum;  // the last thread ends on the chunk total (v = 0 past the end)
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
  if (lane == 0) s_red[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kApThreads / 32; ++w) t += s_red[w];
    out_ap[c] = t;
  }
}

struct ApWorkspace {
  size_t keys_in, keys_out, vals_in, vals_out, class_count, cub_temp, cub_bytes, total;
};
inline ApWorkspace ap_workspace(long long n, int C) {
  ApWorkspace w;
  size_t off = 0;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  w.keys_in = off;     off += ssdx_align_up(nn * 8, 256);
  w.keys_out = off;    off += ssdx_align_up(nn * 8, 256);
  w.vals_in = off;     off += ssdx_align_up(nn * 4, 256);
  w.vals_out = off;    off += ssdx_align_up(nn * 4, 256);
  w.class_count = off; off += ssdx_align_up((size_t)(C + 1) * 4, 256);
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const u64*)nullptr, (u64*)nullptr, (const unsigned*)nullptr,
                                  (unsigned*)nullptr, (int)nn, 0, 64);
  w.cub_bytes = temp;
  w.cub_temp = off;    off += ssdx_align_up(temp, 256);
  w.total = off;
  return w;
}

}  // namespace

extern "C" size_t ssdx_average_precision_workspace_bytes(int64_t num_predictions, int32_t num_classes) {
  if (num_predictions < 0 || num_predictions >= (1ll << 31) || num_classes < 1) return 0;
  return ap_workspace(num_predictions, num_classes).total;
}

extern "C" int ssdx_average_precision(const uint8_t* tp, const float* confidence, const int32_t* label,
                                      int64_t num_predictions, int32_t num_classes, const int64_t* counts,
                                      int32_t background_label, double* out_ap, void* workspace,
                                      size_t workspace_bytes, int device, ssdx_stream stream) {
  SSDX_CHECK_ARG(num_classes >= 1 && num_predictions >= 0 && num_predictions < (1ll << 31),
                 "ssdx_average_precision: bad sizes");
  SSDX_CHECK_ARG(counts != nullptr && out_ap != nullptr, "ssdx_average_precision: NULL pointer argument");
  SSDX_CHECK_ARG(num_predictions == 0 || (tp != nullptr && confidence != nullptr && label != nullptr),
                 "ssdx_average_precision: prediction arrays are NULL");
  const ApWorkspace w = ap_workspace(num_predictions, num_classes);
  if (workspace == nullptr || workspace_bytes < w.total) {
    ssdx_set_error("ssdx_average_precision: workspace too small (%zu < %zu bytes)", workspace_bytes, w.total);
    return SSDX_ERR_WORKSPACE;
  }
  SSDX_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "ssdx_average_precision: workspace must be 256-byte aligned");
  SSDX_DEVICE(device);
  cudaStream_t s = (cudaStream_t)stream;
  char* base = (char*)workspace;
  u64* keys_in = (u64*)(base + w.keys_in);
  u64* keys_out = (u64*)(base + w.keys_out);
  unsigned* vals_in = (unsigned*)(base + w.vals_in);
  unsigned* vals_out = (unsigned*)(base + w.vals_out);
  int* class_count = (int*)(base + w.class_count);
  SSDX_CUDA(cudaMemsetAsync(class_count, 0, (size_t)(num_classes + 1) * 4, s));
  if (num_predictions > 0) {
    const int sms = ssdx_sm_count(device);
    long long grid = (num_predictions + 255) / 256;
    grid = grid > (long long)sms * 8 ? (long long)sms * 8 : grid;
    ssdx_ap_keys_kernel<<<(int)grid, 256, 0, s>>>(tp, confidence, label, (long long)num_predictions, num_classes,
                                                  background_label, keys_in, vals_in, class_count);
    SSDX_LAUNCH_CHECK("ssdx_ap_keys_kernel");
    size_t temp = w.cub_bytes;
    int end_bit = 33;
    while ((1ll << (end_bit - 32)) <= num_classes && end_bit < 64) ++end_bit;  // bits of the class field (0..C)
    SSDX_CUDA(cub::DeviceRadixSort::SortPairs(base + w.cub_temp, temp, keys_in, keys_out, vals_in, vals_out,
                                              (int)num_predictions, 0, end_bit, s));
  } else {
    vals_out = vals_in;
  }
  ssdx_ap_scan_kernel<<<num_classes, kApThreads, 0, s>>>(vals_out, class_count, (const long long*)counts,
                                                          num_classes, background_label, out_ap);
  SSDX_LAUNCH_CHECK("ssdx_ap_scan_kernel");
  return SSDX_OK;
}
