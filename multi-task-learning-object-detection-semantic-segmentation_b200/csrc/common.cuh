// Shared declarations for libssdx (sm_100a).  See include/ssdx.h for the ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ssdx.h"

// One (feature map, box shape) grid of the anchor table: anchor(row, col) = off + (row * W + col) * nb, x extents a
// function of the column only, y extents of the row only (verified bit for bit at creation).  The centre model
// cx(col) ~ cx0 + col / inv_dx, cy(row) ~ cy0 + row / inv_dy and the half extents hx, hy only have to be right up to
// the pads px, py, which hold the measured deviation of the table from the model plus a rounding margin: they size
// the windows of encode_grid.cu, the IoU itself is always evaluated on the table.
struct ssdx_grid_combo {
  int off, H, W, nb;
  float cx0, inv_dx, cy0, inv_dy;
  float hx, hy;      // upper bounds of (x1 - x0) / 2, (y1 - y0) / 2
  float px, py;      // pads on the centre range
  float aw, ah;      // upper bounds of the +1 width / height (datacoder.py:143-144)
  float amin;        // smallest anchor area (datacoder.py:111-114) in the grid
  int ax;            // offset of this grid's axis arrays in ssdx_anchors::axes: x0[W] | x1[W] | y0[H] | y1[H]
};
// A row range of one feature map: anchors [a_lo, a_lo + n), rows [row0, row0 + nrows) of combos [combo0, combo0 + nb)
struct ssdx_grid_span {
  int a_lo, n, combo0, nb, row0, nrows, pad0, pad1;
};
constexpr int kSsdxSpanAnchors = 4096;
constexpr int kSsdxMaxCombos = 256;

// device-resident anchor table (immutable after ssdx_anchors_create)
struct ssdx_anchors {
  int device;
  int32_t A;
  float4* corners;   // (A) x0,y0,x1,y1                       datacoder.py:59-62 / :84-89
  float* area;       // (A) (y1-y0+1)*(x1-x0+1)                datacoder.py:111-114
  float4* enc_cent;  // (A) cx,cy,w,h re-derived from corners  datacoder.py:251-256
  float4* dec_cent;  // (A) cx,cy,w,h as the ctor holds them   datacoder.py:65-70 / :78-81
  // "match order" copy for the encode pair loop: anchors sorted by (area half-octave, descending; Morton code of
  // the centre) and cut into tiles of kSsdxTile, so that one tile is compact in space and in area and whole
  // (tile, GT) pairs can be culled.  The permutation is free: both arg-max tie rules are index-based.
  int32_t n_tiles;
  float4* m_box;   // (n_tiles*kSsdxTile) corners; padding slots hold an empty box (+inf,+inf,-inf,-inf)
  float* m_area;   // (n_tiles*kSsdxTile)
  int* m_idx;      // (n_tiles*kSsdxTile) original anchor index, -1 for padding
  float4* t_bbox;  // (n_tiles) bounding box of the tile's anchors
  float2* t_area;  // (n_tiles) (min, max) anchor area in the tile
  // Grid description (encode_grid.cu), present when the table is a union of regular grids - every output of
  // boxes.DefaultBoundingBoxes is (boxes.py:112-146): per feature map and box shape the x extents depend only on the
  // column and the y extents only on the row.  n_combos == 0: not a grid, the tile match above is used.
  int32_t n_combos;
  ssdx_grid_combo* combos;       // device (n_combos)
  ssdx_grid_combo* combos_host;  // host copy
  int32_t n_spans;
  ssdx_grid_span* spans;         // device (n_spans): row ranges of one map, at most kSsdxSpanAnchors anchors each
  int32_t max_span_anchors;      // largest span
  int32_t max_span_combos;       // largest number of shapes of a map
  // The table itself in separable form: per grid the x extents of its columns and the y extents of its rows, the
  // very floats of the table (corners(row, col) = (x0[col], y0[row], x1[col], y1[row]) bit for bit), so the encode
  // kernels rebuild any anchor, its area and its centroid from a few KB of shared memory instead of gathering them
  // from global memory under a saturated memory system.
  int32_t n_axes;                // floats in axes
  float* axes;                   // device
  int32_t max_map_axes;          // floats of the largest single map's (contiguous) axis arrays
  int2* combo_span;              // device (n_combos): (first span of the grid's map, rows per span)
};

#ifndef SSDX_APT
#define SSDX_APT 2
#endif
constexpr int kSsdxApt = SSDX_APT;        // anchors per lane in one match tile
constexpr int kSsdxTile = 32 * kSsdxApt;  // anchors per match tile

void ssdx_set_error(const char* fmt, ...);

#define SSDX_CHECK_ARG(cond, ...)     \
  do {                                \
    if (!(cond)) {                    \
      ssdx_set_error(__VA_ARGS__);    \
      return SSDX_ERR_INVALID;        \
    }                                 \
  } while (0)

#define SSDX_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      ssdx_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return SSDX_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

// Every entry point runs on the device its handle / `device` argument names and leaves the calling thread's current
// device as it found it (a process may drive several GPUs from one thread).
struct ssdx_device_guard {
  int prev = -1;
  bool changed = false;
  cudaError_t err = cudaSuccess;
  explicit ssdx_device_guard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) {
      err = cudaSetDevice(device);
      changed = (err == cudaSuccess);
    }
  }
  ~ssdx_device_guard() {
    if (changed) cudaSetDevice(prev);
  }
  ssdx_device_guard(const ssdx_device_guard&) = delete;
  ssdx_device_guard& operator=(const ssdx_device_guard&) = delete;
};
#define SSDX_DEVICE(device)            \
  ssdx_device_guard dev_guard__(device); \
  SSDX_CUDA(dev_guard__.err)

// SSDX_DEBUG_SYNC=1 in the environment makes every launch synchronous so a device fault names its kernel
bool ssdx_debug_sync();

#define SSDX_LAUNCH_CHECK(name)                                                     \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ == cudaSuccess && ssdx_debug_sync()) e__ = cudaDeviceSynchronize();     \
    if (e__ != cudaSuccess) {                                                       \
      ssdx_set_error("kernel %s failed: %s", name, cudaGetErrorString(e__));        \
      return SSDX_ERR_CUDA;                                                         \
    }                                                                               \
  } while (0)

static inline size_t ssdx_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Process-level dispatch overrides (ssdx_set_option; default "auto").  All paths of one entry point give identical
// results - the options exist so that tests and tuning runs can pin one.
enum { SSDX_ENC_AUTO = 0, SSDX_ENC_TILE_FUSED, SSDX_ENC_TILE_TILED, SSDX_ENC_GRID_FUSED, SSDX_ENC_GRID_SPLIT };
int ssdx_option_encode_path();
int ssdx_option_nms_cta();  // 0 (auto: from the batch size) | 256 | 512 | 1024 threads per image in the merged NMS pipeline
const char* ssdx_option_nms_path();  // NULL (auto) | "perclass" | "pcta" | "generic" | "blocked"

// number of SMs of a device (cached per device)
int ssdx_sm_count(int device);

// streaming (read-once / write-once) 128-bit accesses: keep the L1 for the anchor table
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// Soft-max of a row of logits, the arithmetic every kernel that accepts logits shares (the Keras soft-max activation
// of the reference's class head: exp(x - max) / sum).  Exponential = ONE hardware instruction, ex2.approx.ftz of
// (x - max) * log2 e (2 ulp + the rounding of the product: ~(2 + 1.2 |x - max|) ulp per term); the sum runs in class
// order; the normaliser is ONE rcp.approx.ftz (1 ulp; the sum lies in [1, C]).  Within ~4e-6 relative of the exact
// value for |x - max| < 25 - inside the 1e-5 fp32 tolerance of the path - and 23 instructions per 4-class row, cheap
// enough to sit inside the NMS score filter.  (expf + IEEE division cost ~140 instructions per row and made that
// kernel ALU-bound: 81 us instead of 33 us per 1024 images; __expf / __frcp_rn without .ftz expand to multi-
// instruction sequences as well.)  One instruction sequence everywhere (-fmad=false: no contraction), so every call
// site produces the same bits.
__device__ __forceinline__ float ssdx_softmax_exp(const float x, const float m) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(__fsub_rn(x, m), 1.4426950408889634f)));
  return e;
}
__device__ __forceinline__ float ssdx_softmax_rcp(const float s) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
  return r;
}
__device__ __forceinline__ float4 ssdx_softmax4(const float4 x) {
  const float m = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
  const float e0 = ssdx_softmax_exp(x.x, m), e1 = ssdx_softmax_exp(x.y, m);
  const float e2 = ssdx_softmax_exp(x.z, m), e3 = ssdx_softmax_exp(x.w, m);
  const float r = ssdx_softmax_rcp(__fadd_rn(__fadd_rn(__fadd_rn(e0, e1), e2), e3));
  return make_float4(__fmul_rn(e0, r), __fmul_rn(e1, r), __fmul_rn(e2, r), __fmul_rn(e3, r));
}

// layers.py:57-79 for one anchor: ((t*s)*d)+c, (exp(t*s)-1)*d, c -/+ (s-1)/2 -> (ymin, xmin, ymax, xmax).  One definition
// shared by the decode kernel and by the NMS kernels that decode their candidates on the fly (ssdx_decode_nms), so
// that both produce the same bits.
__device__ __forceinline__ float4 ssdx_decode_yxyx(const float4 t, const float4 d, const float sx, const float sy,
                                                   const float sw, const float sh) {
  const float cx = __fadd_rn(__fmul_rn(__fmul_rn(t.x, sx), d.z), d.x);
  const float cy = __fadd_rn(__fmul_rn(__fmul_rn(t.y, sy), d.w), d.y);
  const float w = __fmul_rn(__fsub_rn(expf(__fmul_rn(t.z, sw)), 1.0f), d.z);
  const float h = __fmul_rn(__fsub_rn(expf(__fmul_rn(t.w, sh)), 1.0f), d.w);
  const float hw = __fmul_rn(__fsub_rn(w, 1.0f), 0.5f);  // (w-1)/2 exactly
  const float hh = __fmul_rn(__fsub_rn(h, 1.0f), 0.5f);
  return make_float4(__fsub_rn(cy, hh), __fsub_rn(cx, hw), __fadd_rn(cy, hh), __fadd_rn(cx, hw));
}

// Where the NMS kernels get a box from: the decoded (B,A,4) tensor, or the offsets + anchor centroids (decoded per
// candidate, nothing materialised).
struct ssdx_box_src {
  const float4* boxes;    // (B,A) yxyx, or NULL
  const float4* offsets;  // (B,A) standardised centroid offsets when boxes == NULL
  const float4* cent;     // (A) anchor centroids as the decode layer holds them
  float sx, sy, sw, sh;
};
__device__ __forceinline__ float4 ssdx_load_box(const ssdx_box_src& s, const size_t row0, const int idx) {
  if (s.boxes != nullptr) return __ldg(s.boxes + row0 + idx);
  return ssdx_decode_yxyx(__ldg(s.offsets + row0 + idx), __ldg(s.cent + idx), s.sx, s.sy, s.sw, s.sh);
}

// Optional per-phase cycle counters (debug builds only: -DSSDX_PHASE_TIMERS).  Each translation unit that uses
// them defines its own `static __device__ unsigned long long g_phase_cycles[32]` and a reader
// (ssdx_debug_phase_cycles for nms.cu, ssdx_debug_phase_cycles_encode for encode.cu).
#ifdef SSDX_PHASE_TIMERS
#define PHASE_T0() long long pt__ = clock64()
#define PHASE_ADD(slot)                                                        \
  do {                                                                         \
    const long long now__ = clock64();                                         \
    if (threadIdx.x == 0) atomicAdd(&g_phase_cycles[slot], (unsigned long long)(now__ - pt__)); \
    pt__ = now__;                                                              \
  } while (0)
#else
#define PHASE_T0() do {} while (0)
#define PHASE_ADD(slot) do {} while (0)
#endif
