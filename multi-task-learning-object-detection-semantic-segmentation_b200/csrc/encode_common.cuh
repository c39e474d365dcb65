// Device helpers shared by the encode drivers (encode.cu: anchor-tile match for arbitrary anchor tables;
// encode_grid.cu: closed-form windows for anchor tables that are unions of regular grids).
#pragma once

#include "common.cuh"

namespace ssdx_enc {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kFusedMaxG = 256;  // GT boxes per image the one-CTA-per-image drivers stage in shared memory

typedef unsigned long long u64;

// packed (iou bits, inverted anchor index): larger = better IoU, ties -> lower anchor index
__device__ __forceinline__ u64 pack_best(unsigned iou_bits, int a) {
  return ((u64)iou_bits << 32) | (u64)(0xFFFFFFFFu - (unsigned)a);
}
__device__ __forceinline__ int best_anchor(u64 p) { return (int)(0xFFFFFFFFu - (unsigned)(p & 0xFFFFFFFFull)); }
__device__ __forceinline__ unsigned best_bits(const u64* p) { return reinterpret_cast<const unsigned*>(p)[1]; }

__device__ __forceinline__ float gt_area(const float4 b) {
  // datacoder.py:206  (xmax-xmin+1)*(ymax-ymin+1)
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
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

// ---- TMA bulk store (shared -> global), one bulk group per issuing thread ----
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gdst)),
               "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
// the same with an L2 eviction policy (createpolicy.fractional.L2::evict_first / evict_last)
__device__ __forceinline__ void bulk_store_hint(void* gdst, const void* ssrc, unsigned bytes, unsigned long long pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(
                   __cvta_generic_to_global(gdst)),
               "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


}  // namespace ssdx_enc

// encode_grid.cu (host side)
bool ssdx_encode_grid_applies(const ssdx_anchors* anchors, int gmax, int num_classes, float iou_threshold);
bool ssdx_encode_grid_fused_fits(const ssdx_anchors* anchors, int gmax, int num_classes);
int ssdx_encode_grid_launch(const ssdx_anchors* anchors, bool fused, const float* gt_xyxy, const int32_t* gt_label,
                            const int32_t* gt_count, int32_t batch, int32_t gmax, int32_t num_classes,
                            float iou_threshold, const float* std_host, float* out_labels, float* out_offsets,
                            int32_t* out_match, void* workspace, cudaStream_t s);
