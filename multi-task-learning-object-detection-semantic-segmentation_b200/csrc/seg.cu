// Segmentation suppression (reference: layers.py:189-212, wired before the NMS layer in models.py:402-414):
//   class_id  = argmax over the class axis of the (B,H,W,Cm) segmentation output (ties -> lowest index)
//   presence  = clip(sum over batch, height, width of one_hot(class_id, depth), 0, 1)       -> (depth)
//   suppressed = labels_probabilities * presence                                            -> (B,A,C)
// The presence reduction reads the whole mask tensor once (4.9 MB per 480x640x4 image): HBM-bound.  The product is
// either materialised (ssdx_scale_scores) or applied inside the NMS as the scores are read (ssdx_nms_scaled).
#include "common.cuh"

namespace {

constexpr int kSegThreads = 256;
constexpr int kSegBatch = 8;  // independent 128-bit streaming loads in flight per thread

// every CTA ORs the classes it saw into one word; the last CTA to finish converts it to floats and resets the state
__global__ void __launch_bounds__(kSegThreads)
    ssdx_seg_presence_kernel(const float* __restrict__ probs, const long long num_pixels, const int cm, const int depth,
                             unsigned* __restrict__ state /* [0] class bits, [1] finished CTAs */,
                             float* __restrict__ out_presence) {
  const unsigned full = depth >= 32 ? 0xFFFFFFFFu : ((1u << depth) - 1u);
  unsigned seen = 0u;
  const long long stride = (long long)gridDim.x * kSegThreads;
  // presence only ever grows: once every class has been seen (by anybody) the rest of the tensor cannot change it
  const bool done_already = (*reinterpret_cast<volatile unsigned*>(state) & full) == full;
  if (!done_already) {
    if (cm == 4 && (((uintptr_t)probs & 15) == 0)) {
      const float4* p4 = reinterpret_cast<const float4*>(probs);
      // warp-uniform trip count (the bound is tested on the warp's first pixel) so that the vote below is legal
      const int lane = threadIdx.x & 31;
      for (long long w0 = (long long)blockIdx.x * kSegThreads + (threadIdx.x & ~31); w0 < num_pixels;
           w0 += stride * kSegBatch) {
        const long long i0 = w0 + lane;
        float4 v[kSegBatch];
#pragma unroll
        for (int u = 0; u < kSegBatch; ++u) {
          const long long i = i0 + u * stride;
          v[u] = (i < num_pixels) ? ld_stream_f4(p4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kSegBatch; ++u) {
          if (i0 + u * stride < num_pixels) {
            // argmax, first maximum wins
            int best = 0;
            float bv = v[u].x;
            if (v[u].y > bv) { bv = v[u].y; best = 1; }
            if (v[u].z > bv) { bv = v[u].z; best = 2; }
            if (v[u].w > bv) { bv = v[u].w; best = 3; }
            seen |= (best < depth) ? (1u << best) : 0u;  // one_hot of an index >= depth is an all-zero row
          }
        }
        if (__all_sync(0xFFFFFFFFu, (seen & full) == full)) break;  // this warp cannot learn anything new
      }
    } else {
      for (long long i = (long long)blockIdx.x * kSegThreads + threadIdx.x; i < num_pixels; i += stride) {
        const float* p = probs + i * cm;
        int best = 0;
        float bv = __ldg(p);
        for (int c = 1; c < cm; ++c) {
          const float x = __ldg(p + c);
          if (x > bv) { bv = x; best = c; }
        }
        seen |= (best < depth) ? (1u << best) : 0u;
      }
    }
  }
  seen = __reduce_or_sync(0xFFFFFFFFu, seen);
  __shared__ unsigned s_seen;
  __shared__ bool s_last;
  if (threadIdx.x == 0) s_seen = 0u;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && seen) atomicOr(&s_seen, seen);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_seen) atomicOr(&state[0], s_seen);
    __threadfence();
    s_last = (atomicAdd(&state[1], 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    const unsigned bits = *reinterpret_cast<volatile unsigned*>(state);
    for (int c = threadIdx.x; c < depth; c += kSegThreads) out_presence[c] = ((bits >> c) & 1u) ? 1.0f : 0.0f;
    __syncthreads();
    if (threadIdx.x == 0) {  // leave the state clean for the next call on this workspace
      state[0] = 0u;
      state[1] = 0u;
    }
  }
}

template <bool VEC4>
__global__ void __launch_bounds__(256)
    ssdx_scale_scores_kernel(const float* __restrict__ scores, const long long rows, const int C,
                             const float* __restrict__ cscale, float* __restrict__ out) {
  const long long stride = (long long)gridDim.x * 256;
  if (VEC4) {
    const float4 sc = make_float4(__ldg(cscale), __ldg(cscale + 1), __ldg(cscale + 2), __ldg(cscale + 3));
    const float4* in4 = reinterpret_cast<const float4*>(scores);
    float4* out4 = reinterpret_cast<float4*>(out);
    for (long long r = (long long)blockIdx.x * 256 + threadIdx.x; r < rows; r += stride) {
      const float4 v = ld_stream_f4(in4 + r);
      st_stream_f4(out4 + r, make_float4(__fmul_rn(v.x, sc.x), __fmul_rn(v.y, sc.y), __fmul_rn(v.z, sc.z),
                                         __fmul_rn(v.w, sc.w)));
    }
  } else {
    const long long total = rows * C;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += stride)
      out[e] = __fmul_rn(__ldg(scores + e), __ldg(cscale + (int)(e % C)));
  }
}

// probabilities = softmax(logits) [* class_scale], one row per anchor (materialised form; the NMS can also take
// the logits directly: ssdx_nms_logits)
template <bool VEC4>
__global__ void __launch_bounds__(256)
    ssdx_softmax_scores_kernel(const float* __restrict__ logits, const long long rows, const int C,
                               const float* __restrict__ cscale, float* __restrict__ out) {
  const long long stride = (long long)gridDim.x * 256;
  if (VEC4) {
    const float4 sc = cscale ? make_float4(__ldg(cscale), __ldg(cscale + 1), __ldg(cscale + 2), __ldg(cscale + 3))
                             : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4* in4 = reinterpret_cast<const float4*>(logits);
    float4* out4 = reinterpret_cast<float4*>(out);
    for (long long r0 = (long long)blockIdx.x * 256 + threadIdx.x; r0 < rows; r0 += stride * 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long r = r0 + u * stride;
        v[u] = (r < rows) ? ld_stream_f4(in4 + r) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long r = r0 + u * stride;
        if (r < rows) {
          const float4 p = ssdx_softmax4(v[u]);
          st_stream_f4(out4 + r, cscale ? make_float4(__fmul_rn(p.x, sc.x), __fmul_rn(p.y, sc.y), __fmul_rn(p.z, sc.z),
                                                      __fmul_rn(p.w, sc.w))
                                        : p);
        }
      }
    }
  } else {
    for (long long r = (long long)blockIdx.x * 256 + threadIdx.x; r < rows; r += stride) {
      const float* x = logits + r * C;
      float m = __ldg(x);
      for (int c = 1; c < C; ++c) m = fmaxf(m, __ldg(x + c));
      float sum = 0.0f;
      for (int c = 0; c < C; ++c) sum = __fadd_rn(sum, ssdx_softmax_exp(__ldg(x + c), m));
      const float rs = ssdx_softmax_rcp(sum);
      for (int c = 0; c < C; ++c) {
        const float p = __fmul_rn(ssdx_softmax_exp(__ldg(x + c), m), rs);
        out[r * C + c] = cscale ? __fmul_rn(p, __ldg(cscale + c)) : p;
      }
    }
  }
}

}  // namespace

extern "C" int ssdx_softmax_scores(const float* logits, int64_t rows, int32_t num_classes, const float* class_scale,
                                   float* out, int device, ssdx_stream stream) {
  SSDX_CHECK_ARG(logits != nullptr && out != nullptr, "ssdx_softmax_scores: NULL pointer argument");
  SSDX_CHECK_ARG(rows >= 0 && num_classes >= 1, "ssdx_softmax_scores: bad shape");
  if (rows == 0) return SSDX_OK;
  SSDX_DEVICE(device);
  cudaStream_t s = (cudaStream_t)stream;
  const int sms = ssdx_sm_count(device);
  const bool vec4 = num_classes == 4 && (((uintptr_t)logits | (uintptr_t)out) & 15) == 0;
  long long grid = (rows + 256 * 4 - 1) / (256 * 4);
  const long long cap = (long long)sms * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  if (vec4)
    ssdx_softmax_scores_kernel<true><<<(int)grid, 256, 0, s>>>(logits, rows, num_classes, class_scale, out);
  else
    ssdx_softmax_scores_kernel<false><<<(int)grid, 256, 0, s>>>(logits, rows, num_classes, class_scale, out);
  SSDX_LAUNCH_CHECK("ssdx_softmax_scores_kernel");
  return SSDX_OK;
}

extern "C" int ssdx_seg_class_presence(const float* seg_probs, int64_t num_pixels, int32_t cm, int32_t depth,
                                       float* out_presence, void* workspace, size_t workspace_bytes, int device,
                                       ssdx_stream stream) {
  SSDX_CHECK_ARG(seg_probs != nullptr && out_presence != nullptr, "ssdx_seg_class_presence: NULL pointer argument");
  SSDX_CHECK_ARG(num_pixels >= 1 && cm >= 1, "ssdx_seg_class_presence: bad shape (pixels=%lld, classes=%d)",
                 (long long)num_pixels, cm);
  SSDX_CHECK_ARG(depth >= 1 && depth <= 32, "ssdx_seg_class_presence: depth must be in [1, 32] (got %d)", depth);
  if (workspace == nullptr || workspace_bytes < 256) {
    ssdx_set_error("ssdx_seg_class_presence: workspace too small (%zu < 256 bytes)", workspace_bytes);
    return SSDX_ERR_WORKSPACE;
  }
  SSDX_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "ssdx_seg_class_presence: workspace must be 256-byte aligned");
  SSDX_DEVICE(device);
  cudaStream_t s = (cudaStream_t)stream;
  SSDX_CUDA(cudaMemsetAsync(workspace, 0, 8, s));
  const int sms = ssdx_sm_count(device);
  const long long per_cta = (long long)kSegThreads * kSegBatch;
  long long grid = (num_pixels + per_cta - 1) / per_cta;
  grid = grid < 1 ? 1 : (grid > (long long)sms * 8 ? (long long)sms * 8 : grid);
  ssdx_seg_presence_kernel<<<(int)grid, kSegThreads, 0, s>>>(seg_probs, (long long)num_pixels, cm, depth,
                                                             (unsigned*)workspace, out_presence);
  SSDX_LAUNCH_CHECK("ssdx_seg_presence_kernel");
  return SSDX_OK;
}

extern "C" int ssdx_scale_scores(const float* scores, int64_t rows, int32_t num_classes, const float* class_scale,
                                 float* out, int device, ssdx_stream stream) {
  SSDX_CHECK_ARG(scores != nullptr && class_scale != nullptr && out != nullptr, "ssdx_scale_scores: NULL pointer argument");
  SSDX_CHECK_ARG(rows >= 0 && num_classes >= 1, "ssdx_scale_scores: bad shape");
  if (rows == 0) return SSDX_OK;
  SSDX_DEVICE(device);
  cudaStream_t s = (cudaStream_t)stream;
  const int sms = ssdx_sm_count(device);
  const bool vec4 = num_classes == 4 && (((uintptr_t)scores | (uintptr_t)out) & 15) == 0;
  const long long work = vec4 ? rows : rows * num_classes;
  long long grid = (work + 256 * 4 - 1) / (256 * 4);
  grid = grid < 1 ? 1 : (grid > (long long)sms * 16 ? (long long)sms * 16 : grid);
  if (vec4)
    ssdx_scale_scores_kernel<true><<<(int)grid, 256, 0, s>>>(scores, rows, num_classes, class_scale, out);
  else
    ssdx_scale_scores_kernel<false><<<(int)grid, 256, 0, s>>>(scores, rows, num_classes, class_scale, out);
  SSDX_LAUNCH_CHECK("ssdx_scale_scores_kernel");
  return SSDX_OK;
}
