// SSD target encode for anchor tables that are unions of regular grids (reference: datacoder.py:205-300).
//
// Every table boxes.DefaultBoundingBoxes produces is one: per feature map and box shape the anchors are a grid
// whose x extents depend only on the column and whose y extents depend only on the row (boxes.py:112-146;
// ssdx_anchors_create verifies that bit for bit and falls back to the tile match of encode.cu for any other table).
// Same parity contract as encode.cu - the IoU of a pair is the reference's fp32 expression tree evaluated on the
// anchor table, both arg-max ties go to the lower index, duplicates resolve by the closed form - but the pairs that
// are evaluated at all are found in closed form instead of by search:
//
//   a pair (a, g) matters iff  iou > thr  (anchor side, :236-241)  or  iou >= max_a iou[:, g]  (GT side, :230-231),
//   hence only if  iou >= t(g) = min(max(thr, 0), seed(g))  for ANY lower bound seed(g) of the GT's best IoU;
//   iou >= t  =>  inter >= t * (area_a + area_g) / (1 + t) =: Imin, and  inter = iw * ih  with iw <= min(aw, gw),
//   ih <= min(ah, gh)  =>  iw >= Imin / min(ah, gh)  and  ih >= Imin / min(aw, gw);
//   iw as a function of the anchor centre is min(cx + hx, gx1) - max(cx - hx, gx0) + 1, so iw >= n confines cx to
//   [gx0 + n - 1 - hx, gx1 - n + 1 + hx]: a column range of the grid, and likewise a row range.
//
// Per (GT box, grid) that is one rectangle of cells ("window"); all bounds are padded by the measured deviation of
// the table from its linear centre model plus a rounding margin (ssdx_grid_combo::px/py), so a window can only be too
// large, and every cell inside is decided by the exact IoU.  seed(g) is the exact IoU of the anchor nearest to the GT
// centre in every grid.  On the reference configuration (9600 anchors, 16 GT boxes) the windows hold ~250 cells per
// image for ~200 pairs above the threshold - the 153,600-pair matrix is never looked at.
//
// The anchors themselves are rebuilt from per-grid axis arrays x0[W] | x1[W] | y0[H] | y1[H] (the very floats of the
// table) held in shared memory: corners, area and centroid of any cell, bit for bit, without a global read.
//
// Drivers:
//   fused : persistent, warp-specialised CTAs, one image at a time per CTA (tables up to 32768 anchors / 256 GT
//           boxes).  A producer warp streams the constant background rows of the image as TMA bulk stores, a few in
//           flight at a time; the matching warps own whole GT boxes (lanes = grids: seed, bound, windows and their
//           prefix sum stay inside the warp), the few anchors with a pair above the threshold claim slots of a
//           compact table (pass 0: slot max IoU + per-GT best; pass 1: lowest GT index among a slot's maxima), then
//           the closed form, and the positives are patched over the background as whole 32-byte sectors.
//   split : K1 grid (32 GT boxes, image) walks the windows once: per-GT best pairs + a hit list per row span of a
//           feature map; K2 grid (span, image) reduces its list, applies the closed form and writes the span's rows:
//           larger tables (dense stress: 100k anchors x 256 GT boxes).
#include <algorithm>

#include "common.cuh"
#include "encode_common.cuh"
#include "grid_window.cuh"

#ifdef SSDX_PHASE_TIMERS
static __device__ unsigned long long g_phase_cycles[32];
extern "C" __attribute__((visibility("default"))) int ssdx_debug_phase_cycles_encode_grid(unsigned long long* out32) {
  unsigned long long zero[32] = {0};
  if (cudaMemcpyFromSymbol(out32, g_phase_cycles, sizeof(zero)) != cudaSuccess) return -2;
  if (cudaMemcpyToSymbol(g_phase_cycles, zero, sizeof(zero)) != cudaSuccess) return -2;
  return 0;
}
#endif

namespace {

using namespace ssdx_enc;

#ifndef SSDX_GRID_MINB
#define SSDX_GRID_MINB 4
#endif
#ifndef SSDX_GRID_L2HINT
#define SSDX_GRID_L2HINT 1  // 1: label rows leave with L2 evict_first (nothing on the device re-reads them soon; the offsets
                            // that decode reads next then stay in L2: decode 61 -> 55 us in the step); 2: and offset rows
                            // with evict_last (measured: no further gain)
#endif
#ifndef SSDX_GRID_BGROWS
#define SSDX_GRID_BGROWS 256
#endif
constexpr int kThreads = 256;
constexpr int kWinCap = 512;                 // windows per round (GT boxes of the round x grids)
constexpr int kBgRows = SSDX_GRID_BGROWS;    // background rows per bulk store
constexpr int kBestG = 32;                   // GT boxes per CTA of the split driver's K1
constexpr int kGtTile = 256;                 // GT boxes staged per round by the split driver's K2
constexpr unsigned short kNone = 0xFFFFu;
constexpr int kGridFusedMaxA = 32768;        // anchor -> slot table of the fused driver: 2 bytes per anchor

// corners of anchor (row, col) of a grid, the floats of the table (ssdx_anchors::axes)
__device__ __forceinline__ float4 grid_anchor_box(const ssdx_grid_combo& cb, const float* ax, const int r,
                                                  const int c) {
  const float* p = ax + cb.ax;
  return make_float4(p[c], p[2 * cb.W + r], p[cb.W + c], p[2 * cb.W + cb.H + r]);
}
// datacoder.py:111-114, rounded as ssdx_anchors_create rounds the area table
__device__ __forceinline__ float anchor_area(const float4 b) {
  return __fmul_rn(__fadd_rn(__fsub_rn(b.w, b.y), 1.0f), __fadd_rn(__fsub_rn(b.z, b.x), 1.0f));
}
// datacoder.py:141-144 (cx, cy, w, h), rounded as ssdx_anchors_create rounds enc_cent
__device__ __forceinline__ float4 anchor_centroid(const float4 b) {
  return make_float4(__fmul_rn(__fadd_rn(b.z, b.x), 0.5f), __fmul_rn(__fadd_rn(b.w, b.y), 0.5f),
                     __fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
}

// datacoder.py:210-220 for one pair; 0 when the boxes do not intersect (the reference's max(0, .) makes inter 0)
__device__ __forceinline__ float pair_iou(const float4 ab, const float aa, const float4 gb, const float ga) {
  const float iw = __fadd_rn(__fsub_rn(fminf(ab.z, gb.z), fmaxf(ab.x, gb.x)), 1.0f);
  const float ih = __fadd_rn(__fsub_rn(fminf(ab.w, gb.w), fmaxf(ab.y, gb.y)), 1.0f);
  if (!(iw > 0.0f && ih > 0.0f)) return 0.0f;
  const float inter = __fmul_rn(iw, ih);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ga), inter));
}

// seed of the GT-side bound: exact IoU of the anchor of this grid nearest to the GT centre
__device__ __forceinline__ void grid_seed(const ssdx_grid_combo& cb, const float* ax, const float4 gb,
                                          const float ga, u64* s_best_g) {
  const float gw = gb.z - gb.x + 1.0f, gh = gb.w - gb.y + 1.0f;
  if (!(gw > 0.0f && gh > 0.0f)) return;
  const float fc = rintf((0.5f * (gb.x + gb.z) - cb.cx0) * cb.inv_dx);
  const float fr = rintf((0.5f * (gb.y + gb.w) - cb.cy0) * cb.inv_dy);
  const int c = (int)fminf(fmaxf(fc, 0.0f), (float)(cb.W - 1));
  const int r = (int)fminf(fmaxf(fr, 0.0f), (float)(cb.H - 1));
  const float4 ab = grid_anchor_box(cb, ax, r, c);
  const float iou = pair_iou(ab, anchor_area(ab), gb, ga);
  if (iou > 0.0f) atomicMax(s_best_g, pack_best(__float_as_uint(iou), cb.off + (r * cb.W + c) * cb.nb));
}

// (grid, row, col) of anchor a: the grids of one feature map are consecutive and share H, W, nb
__device__ __forceinline__ float4 anchor_box_by_index(const ssdx_grid_combo* s_combo, const int nq, const float* ax,
                                                      const int a) {
  int q = 0;
  while (q + s_combo[q].nb < nq && a >= s_combo[q + s_combo[q].nb].off) q += s_combo[q].nb;  // first grid of the map
  const ssdx_grid_combo& m = s_combo[q];
  const int rel = a - m.off;
  const int cell = rel / m.nb, k = rel - cell * m.nb;
  const int r = cell / m.W, c = cell - r * m.W;
  return grid_anchor_box(s_combo[q + k], ax, r, c);
}

__device__ __forceinline__ void atomic_min_u16(unsigned short* p, const unsigned short v) {
  unsigned short old = *p;
  while (v < old) {
    const unsigned short prev = atomicCAS(p, old, v);
    if (prev == old) break;
    old = prev;
  }
}
__device__ __forceinline__ void atomic_max_u16(unsigned short* p, const unsigned short v) {
  unsigned short old = *p;
  while (v > old) {
    const unsigned short prev = atomicCAS(p, old, v);
    if (prev == old) break;
    old = prev;
  }
}

// ------------------------------------------------------------------------------------------ fused driver
// One CTA per image.  Each warp owns whole GT boxes (g = warp, warp + kWarps, ...): its lanes are the grids, so the
// seed, the bound t(g), the windows and their prefix sum stay inside the warp (shuffles, no barrier, no shared
// scratch), the cells of a GT box are walked 32 at a time, and the GT-side arg-max is a warp reduction.  What crosses
// warps is the anchor side: an anchor reached by windows of several GT boxes.  Anchors with a pair above the
// threshold are rare (~2 %), so they get a slot in a compact table (s_slot: anchor -> slot, claimed with one 16-bit
// CAS; per slot the max IoU bits, the cell, the lowest GT index among the maxima) instead of per-anchor tables - the
// slots double as the list of positives, and the CTA fits four times on an SM.
//   pass 0: per-slot max IoU, per-GT best anchor | barrier | pass 1: lowest GT index among a slot's maxima | barrier
//   | closed form (may add slots: anchors that are only some GT's best) | barrier | encode + patch the slots' rows.
// An image with more positives than slots (kSlots; degenerate inputs) is redone by the exact per-anchor loop
// grid_fused_overflow.
constexpr int kSlots = 1024;
constexpr unsigned kCellDead = 0xFFFFFFFFu;
// Warp specialisation: kCompWarps warps match, one more warp only feeds the TMA unit with the background rows, a few
// bulk stores in flight at a time.  (Issued all at once - 75 bulk stores per CTA, four CTAs per SM - the copies back
// up in the SM's memory-instruction queue and every shared-memory access of the matching warps waits behind them:
// 74 us per 1024 images, the sum of the match (29 us) and the output stream (45 us) instead of their maximum.)
#ifndef SSDX_GRID_COMP_WARPS
#define SSDX_GRID_COMP_WARPS 8
#endif
#ifndef SSDX_GRID_INFLIGHT
#define SSDX_GRID_INFLIGHT 6
#endif
constexpr int kCompWarps = SSDX_GRID_COMP_WARPS;
constexpr int kCompThreads = kCompWarps * 32;
constexpr int kFThreads = kCompThreads + 32;
constexpr int kBulkInFlight = SSDX_GRID_INFLIGHT;  // bulk stores of one CTA whose source has not been read yet

__device__ __forceinline__ void comp_barrier() {  // the matching warps only (named barrier 1)
  asm volatile("bar.sync 1, %0;" ::"n"(kCompThreads) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

__host__ __device__ inline int grid_label_rows(int C) {  // rows per constant label block (multiple of 4)
  int r = ((kBgRows * 16) / (4 * C)) & ~3;
  return r < 4 ? 4 : r;
}
__host__ __device__ inline bool grid_bulk_labels(int A, int C) { return ((A * (long long)C) % 4) == 0 && C <= 1024; }

__host__ __device__ inline size_t grid_fused_smem_bytes(int A, int G, int C, int nq, int n_axes) {
  const size_t g = (size_t)(G > 0 ? G : 1);
  size_t n = (size_t)kBgRows * 16;                                                  // bg_off
  n += grid_bulk_labels(A, C) ? (size_t)grid_label_rows(C) * C * 4 : 16;            // bg_lab
  n += g * (16 + 8 + 4 + 4 + 4 + 4);                                                // box, best, area, lab, bcell, s1a
  n = (n + 15) / 16 * 16;
  n += (size_t)nq * sizeof(ssdx_grid_combo) + (size_t)n_axes * 4;                   // grids, axis arrays
  n += (size_t)kWinCap * 8;                                                         // window rectangles
  n += (size_t)kSlots * (4 + 4 + 2) + 64;                                           // slots, ctrl
  n += (size_t)((A + 7) / 8 * 8) * 2;                                               // anchor -> slot
  return n + 16;
}

__device__ __forceinline__ float4 one_hot4(const int lab) {  // tf.one_hot(lab, 4): all-zero row outside [0, 4)
  return make_float4(lab == 0 ? 1.f : 0.f, lab == 1 ? 1.f : 0.f, lab == 2 ? 1.f : 0.f, lab == 3 ? 1.f : 0.f);
}
// two adjacent 16-byte rows = one 32-byte sector, p 32-byte aligned: one 256-bit store
__device__ __forceinline__ void st_stream_sector(float4* p, const float4 lo, const float4 hi) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(lo.x), "f"(lo.y), "f"(lo.z),
               "f"(lo.w), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w)
               : "memory");
}

// cell code of (grid q, row r, col c): q < 256, r, c < 4096 (ssdx_anchors_create keeps grids within that)
__device__ __forceinline__ unsigned cell_code(const int q, const int r, const int c) {
  return ((unsigned)q << 24) | ((unsigned)r << 12) | (unsigned)c;
}

// slot of anchor a, claimed on first use; -1 when the table is full
__device__ __forceinline__ int get_slot(unsigned short* s_slot, int* s_nslots, const int cap, const int a) {
  const unsigned short cur = *(volatile unsigned short*)(s_slot + a);
  if (cur != kNone) return cur;
  const int mine = atomicAdd(s_nslots, 1);
  if (mine >= cap) return -1;
  const unsigned short prev = atomicCAS(s_slot + a, kNone, (unsigned short)mine);
  return prev == kNone ? mine : (int)prev;  // lost the race: slot `mine` stays dead (its cell keeps kCellDead)
}

// Walk the cells of up to 32 windows held one per lane (rectangle rx = r0 | c0 << 16, ry = columns | rows << 16,
// ry = 0: empty; tag = anything the owner wants passed along): f(tag of the owning lane, row, col) for every cell,
// 32 cells per round.  All lanes of the warp must call it.
template <typename F>
__device__ __forceinline__ void warp_for_each_cell_tagged(const int rx, const int ry, const int tag, const int lane,
                                                          F f) {
  const int cnt = (ry & 0xFFFF) * (ry >> 16);
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int u = __shfl_up_sync(kFull, incl, d);
    if (lane >= d) incl += u;
  }
  const int total = __shfl_sync(kFull, incl, 31);
  const int excl = incl - cnt;
  for (int i0 = 0; i0 < total; i0 += 32) {
    const int i = i0 + lane;
    int lo = 0;  // largest lane l with excl(l) <= i: the window cell i belongs to (empty windows share a start)
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const int e = __shfl_sync(kFull, excl, (lo + step) & 31);
      if (e <= i) lo += step;
    }
    const int oe = __shfl_sync(kFull, excl, lo);
    const int ox = __shfl_sync(kFull, rx, lo);
    const int oy = __shfl_sync(kFull, ry, lo) & 0xFFFF;
    const int ot = __shfl_sync(kFull, tag, lo);
    if (i < total) {
      const int local = i - oe;
      const int rr = local / oy, cc = local - rr * oy;
      f(ot, (ox & 0xFFFF) + rr, (ox >> 16) + cc);
    }
  }
}
template <typename F>
__device__ __forceinline__ void warp_for_each_cell(const int rx, const int ry, const int lane, F f) {
  warp_for_each_cell_tagged(rx, ry, lane, lane, f);  // tag = owning lane
}

// The exact per-anchor loop, for an image that does not fit the slot table: every anchor against every GT box.
// Reuses s_slot as the anchor -> matched GT table.
__device__ __noinline__ void grid_fused_overflow(const ssdx_grid_combo* s_combo, const int nq, const float* s_ax,
                                                 const float4* s_box, const float* s_area, u64* s_best,
                                                 int* s_s1a, unsigned short* s_tab, const int A, const int G,
                                                 const float iou_thr) {
  const int tid = threadIdx.x;
  const int A8 = (A + 7) / 8 * 8;
  comp_barrier();
  for (int i = tid; i < A8 / 8; i += kCompThreads) reinterpret_cast<int4*>(s_tab)[i] = make_int4(-1, -1, -1, -1);
  for (int g = tid; g < G; g += kCompThreads) s_best[g] = 0ull;
  comp_barrier();
  for (int a = tid; a < A; a += kCompThreads) {
    const float4 ab = anchor_box_by_index(s_combo, nq, s_ax, a);
    const float aa = anchor_area(ab);
    float bv = 0.0f;
    int bg = 0;
    for (int g = 0; g < G; ++g) {
      const float iou = pair_iou(ab, aa, s_box[g], s_area[g]);
      if (iou > bv) {  // strict: ties keep the lower GT index
        bv = iou;
        bg = g;
      }
      const unsigned bits = __float_as_uint(iou);
      if (iou > 0.0f && bits >= best_bits(s_best + g)) atomicMax(s_best + g, pack_best(bits, a));
    }
    if (bv > iou_thr) s_tab[a] = (unsigned short)bg;
  }
  comp_barrier();
  for (int g = tid; g < G; g += kCompThreads) {
    const u64 p = s_best[g];
    s_s1a[g] = ((p >> 32) != 0ull) ? best_anchor(p) : -1;
  }
  comp_barrier();
  for (int g = tid; g < G; g += kCompThreads) {
    const int a = s_s1a[g];
    if (a < 0) continue;
    bool is_max = true;
    for (int h = g + 1; h < G; ++h) is_max = is_max && (s_s1a[h] != a);
    if (is_max) {
      const unsigned short g2 = s_tab[a];
      s_tab[a] = (g2 != kNone && s_s1a[g2] != a) ? g2 : (unsigned short)g;
    }
  }
  comp_barrier();
}

// Persistent CTAs: grid = resident CTAs of the device, CTA c takes images c, c + gridDim.x, ...  (Measured against a
// ticket counter zeroed by a memset node before the launch: 62.8 us static, 66.2 us with tickets per 1024 images -
// every image writes the same number of bytes, so there is little to balance and the extra stream operation costs
// more than it saves.)  The producer warp keeps the output stream of the CTA running across images.  Per image the
// two sides meet once (named barrier 2), when the image's background rows are complete and its match is done; the
// matching warps then patch the positives and go on.
__global__ void __launch_bounds__(kFThreads, SSDX_GRID_MINB)
    ssdx_encode_grid_fused_kernel(const ssdx_grid_combo* __restrict__ combos, const int nq,
                                  const float* __restrict__ axes, const int n_axes,
                                  const float4* __restrict__ gt_xyxy, const int* __restrict__ gt_label,
                                  const int* __restrict__ gt_count, const int B, const int A, const int Gmax,
                                  const int C, const float iou_thr, const float sx, const float sy, const float sw,
                                  const float sh, float* __restrict__ out_labels, float4* __restrict__ out_offsets,
                                  int* __restrict__ out_match) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Gs = Gmax > 0 ? Gmax : 1;
  const bool bulk_lab = grid_bulk_labels(A, C);
  const int lab_rows = grid_label_rows(C);
  const int A8 = (A + 7) / 8 * 8;

  float4* bg_off;
  float* bg_lab;
  float4* s_box;
  u64* s_best;
  float* s_area;
  int* s_lab;
  unsigned* s_bcell;
  int* s_s1a;
  ssdx_grid_combo* s_combo;
  float* s_ax;
  int2* s_rect;
  unsigned* s_cval;
  unsigned* s_ccell;
  unsigned short* s_cg;
  int* s_ctrl;
  unsigned short* s_slot;
  {
    unsigned char* p = smem_raw;
    bg_off = (float4*)p;  p += (size_t)kBgRows * 16;
    bg_lab = (float*)p;   p += bulk_lab ? (size_t)lab_rows * C * 4 : 16;
    s_box = (float4*)p;   p += (size_t)Gs * 16;
    s_best = (u64*)p;     p += (size_t)Gs * 8;
    s_area = (float*)p;   p += (size_t)Gs * 4;
    s_lab = (int*)p;      p += (size_t)Gs * 4;
    s_bcell = (unsigned*)p; p += (size_t)Gs * 4;
    s_s1a = (int*)p;      p += (size_t)Gs * 4;
    p = smem_raw + ((size_t)(p - smem_raw) + 15) / 16 * 16;
    s_combo = (ssdx_grid_combo*)p; p += (size_t)nq * sizeof(ssdx_grid_combo);
    s_ax = (float*)p;     p += (size_t)n_axes * 4;
    s_rect = (int2*)p;    p += (size_t)kWinCap * 8;
    s_cval = (unsigned*)p;  p += (size_t)kSlots * 4;
    s_ccell = (unsigned*)p; p += (size_t)kSlots * 4;
    s_cg = (unsigned short*)p; p += (size_t)kSlots * 2;
    s_ctrl = (int*)p;     p += 64;
    s_slot = (unsigned short*)p;
  }
  int* s_nslots = s_ctrl + 0;
  int* s_overflow = s_ctrl + 1;

  // ---- once per CTA: grids, axis arrays (the table in separable form), constant background blocks, empty tables ----
  PHASE_T0();
  for (int i = tid; i < nq * (int)(sizeof(ssdx_grid_combo) / 16); i += kFThreads)
    reinterpret_cast<int4*>(s_combo)[i] = __ldg(reinterpret_cast<const int4*>(combos) + i);
  for (int i = tid; i < n_axes / 4; i += kFThreads)
    reinterpret_cast<int4*>(s_ax)[i] = __ldg(reinterpret_cast<const int4*>(axes) + i);
  for (int i = tid; i < kBgRows; i += kFThreads) bg_off[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bulk_lab) {
    if (C == 4) {
      for (int r = tid; r < lab_rows; r += kFThreads)
        reinterpret_cast<float4*>(bg_lab)[r] = make_float4(1.f, 0.f, 0.f, 0.f);
    } else {
      for (int e = tid; e < lab_rows * C; e += kFThreads) bg_lab[e] = (e % C == 0) ? 1.0f : 0.0f;
    }
  }
  for (int i = tid; i < kSlots; i += kFThreads) {
    s_cval[i] = 0u;
    s_ccell[i] = kCellDead;
    s_cg[i] = kNone;
  }
  for (int i = tid; i < A8 / 8; i += kFThreads) reinterpret_cast<int4*>(s_slot)[i] = make_int4(-1, -1, -1, -1);
  if (tid == 0) {
    *s_nslots = 0;
    *s_overflow = 0;
  }
  fence_async_smem();  // the background blocks are read by the async proxy
  __syncthreads();
  PHASE_ADD(0);

  if (warp == kCompWarps) {
    // =================== producer warp: background rows of every image of this CTA as TMA bulk stores ===================
    const int n_off = (A + kBgRows - 1) / kBgRows;
    const int n_lab = bulk_lab ? (A + lab_rows - 1) / lab_rows : 0;
    const int n_st = n_off + n_lab;
#if SSDX_GRID_L2HINT >= 1
    const unsigned long long pol_first = l2_policy_evict_first();
#endif
#if SSDX_GRID_L2HINT >= 2
    const unsigned long long pol_last = l2_policy_evict_last();
#endif
    auto issue = [&](const int img, const int i0, const int i1) {  // bulk stores [i0, i1) of an image, one group each
      float4* offs = out_offsets + (size_t)img * A;
      float* labs = out_labels + (size_t)img * A * C;
      for (int i = i0; i < i1; ++i) {
        if (i < n_off) {
          const int r0 = i * kBgRows;
#if SSDX_GRID_L2HINT >= 2
          bulk_store_hint(offs + r0, bg_off, (unsigned)(min(kBgRows, A - r0) * 16), pol_last);
#else
          bulk_store(offs + r0, bg_off, (unsigned)(min(kBgRows, A - r0) * 16));
#endif
        } else {
          const int r0 = (i - n_off) * lab_rows;
#if SSDX_GRID_L2HINT >= 1
          bulk_store_hint(labs + (size_t)r0 * C, bg_lab, (unsigned)(min(lab_rows, A - r0) * C * 4), pol_first);
#else
          bulk_store(labs + (size_t)r0 * C, bg_lab, (unsigned)(min(lab_rows, A - r0) * C * 4));
#endif
        }
        bulk_commit();
        bulk_wait_read<kBulkInFlight>();  // at most kBulkInFlight copies of this CTA queued in the SM
      }
    };
    const bool lookahead = n_st >= kBulkInFlight;  // the first stores of the next image go out before this one is waited for
    int img = blockIdx.x;
    int head = 0;  // stores of img already issued
    while (img < B) {
      const int next = img + gridDim.x;
#ifndef SSDX_GRID_EXP_NOBULK
      if (lane == 0) {
        issue(img, head, n_st);
        if (next < B && lookahead) {
          issue(next, 0, kBulkInFlight);
          bulk_wait<kBulkInFlight>();  // everything but the kBulkInFlight newest groups: image img is complete
        } else {
          bulk_wait_all();
        }
      }
#endif
      head = (next < B && lookahead) ? kBulkInFlight : 0;
      __syncwarp();
      asm volatile("bar.sync 2, %0;" ::"n"(kFThreads) : "memory");  // image img: background rows in place
      img = next;
    }
  } else {
    // =================== matching warps ===================
    const float thr_eff = fmaxf(iou_thr, 0.0f);
    int img = blockIdx.x;
    while (img < B) {
      const int b = img;
      int G = gt_count[b];
      G = G < 0 ? 0 : (G > Gmax ? Gmax : G);
      float4* offs = out_offsets + (size_t)b * A;
      float* labs = out_labels + (size_t)b * A * C;
      for (int g = tid; g < G; g += kCompThreads) {
        const float4 bx = __ldg(gt_xyxy + (size_t)b * Gmax + g);
        s_box[g] = bx;
        s_area[g] = gt_area(bx);
        s_lab[g] = __ldg(gt_label + (size_t)b * Gmax + g);
      }
      if (out_match)  // (diagnostic output) background everywhere first, the positives overwrite it after a barrier
        for (int a = tid; a < A; a += kCompThreads) out_match[(size_t)b * A + a] = -1;
      comp_barrier();
      PHASE_ADD(6);
      const int slot_cap = kSlots - G;  // the closed form may add up to G positives of its own (G <= kFusedMaxG)
      // ---- pass 0: seeds, windows, cells: per-slot max IoU and the GT side ----
      for (int g = warp; g < G; g += kCompWarps) {
        const float4 gb = s_box[g];
        const float ga = s_area[g];
        u64 best = 0ull;       // this lane's best pair of GT box g: (IoU bits, ~anchor)
        unsigned bcell = 0u;   // and its cell
        // seeds: the anchor of every grid nearest to the GT centre
        const float gw = gb.z - gb.x + 1.0f, gh = gb.w - gb.y + 1.0f;
        const bool gt_ok = gw > 0.0f && gh > 0.0f;  // an empty or NaN box: every IoU is 0 / NaN and never wins a '>'
        if (gt_ok)
          for (int q = lane; q < nq; q += 32) {
            const ssdx_grid_combo& cb = s_combo[q];
            const float fc = rintf((0.5f * (gb.x + gb.z) - cb.cx0) * cb.inv_dx);
            const float fr = rintf((0.5f * (gb.y + gb.w) - cb.cy0) * cb.inv_dy);
            const int c = (int)fminf(fmaxf(fc, 0.0f), (float)(cb.W - 1));
            const int r = (int)fminf(fmaxf(fr, 0.0f), (float)(cb.H - 1));
            const float4 ab = grid_anchor_box(cb, s_ax, r, c);
            const float iou = pair_iou(ab, anchor_area(ab), gb, ga);
            const u64 p = pack_best(__float_as_uint(iou), cb.off + (r * cb.W + c) * cb.nb);
            if (iou > 0.0f && p > best) {
              best = p;
              bcell = cell_code(q, r, c);
            }
          }
        const float seed = __uint_as_float(__reduce_max_sync(kFull, (unsigned)(best >> 32)));
        const float t = fminf(thr_eff, seed);
        for (int q0 = 0; q0 < nq; q0 += 32) {
          const int q = q0 + lane;
          int rx = 0, ry = 0;
          ssdx_grid_rect rc;
          if (q < nq && gt_ok && ssdx_grid_window(s_combo[q], gb, t, 0, s_combo[q].H, rc)) {
            rx = rc.r0 | (rc.c0 << 16);
            ry = rc.nc | (rc.nr << 16);
          }
          if (g * nq + q < kWinCap && q < nq) s_rect[g * nq + q] = make_int2(rx, ry);  // pass 1 walks it again
          warp_for_each_cell(rx, ry, lane, [&](const int ql, const int r, const int c) {
            const ssdx_grid_combo& cb = s_combo[q0 + ql];
            const int a = cb.off + (r * cb.W + c) * cb.nb;
            const float4 ab = grid_anchor_box(cb, s_ax, r, c);
            const float iou = pair_iou(ab, anchor_area(ab), gb, ga);
            const unsigned bits = __float_as_uint(iou);
            if (iou > iou_thr) {
              const int slot = get_slot(s_slot, s_nslots, slot_cap, a);
              if (slot >= 0) {
                atomicMax(s_cval + slot, bits);
                s_ccell[slot] = cell_code(q0 + ql, r, c);  // (every writer of a slot writes the same cell)
              } else {
                *s_overflow = 1;
              }
            }
            const u64 p = pack_best(bits, a);
            if (iou > 0.0f && p > best) {
              best = p;
              bcell = cell_code(q0 + ql, r, c);
            }
          });
        }
        // GT side: arg-max over the lanes, ties -> lowest anchor index (largest ~anchor)
        const unsigned hi = __reduce_max_sync(kFull, (unsigned)(best >> 32));
        const unsigned lo = __reduce_max_sync(kFull, (unsigned)(best >> 32) == hi ? (unsigned)best : 0u);
        if ((unsigned)(best >> 32) == hi && (unsigned)best == lo) {  // one lane (distinct anchors), or all when hi == 0
          s_best[g] = ((u64)hi << 32) | lo;
          s_bcell[g] = bcell;
        }
      }
      comp_barrier();
      PHASE_ADD(1);

      // ---- pass 1: lowest GT index among the maxima of a slot (walks the windows of pass 0 again) ----
      if (*s_overflow == 0) {
        for (int g = warp; g < G; g += kCompWarps) {
          const float4 gb = s_box[g];
          const float ga = s_area[g];
          const float gw = gb.z - gb.x + 1.0f, gh = gb.w - gb.y + 1.0f;
          const bool gt_ok = gw > 0.0f && gh > 0.0f;
          float t = 0.0f;
          if (G * nq > kWinCap)  // windows beyond the stored ones are rebuilt; the bound may only have grown since
            t = fminf(thr_eff, __uint_as_float(best_bits(s_best + g)));  // pass 0, the pairs above thr stay inside
          for (int q0 = 0; q0 < nq; q0 += 32) {
            const int q = q0 + lane;
            int rx = 0, ry = 0;
            if (q < nq) {
              if (g * nq + q < kWinCap) {
                const int2 w = s_rect[g * nq + q];
                rx = w.x;
                ry = w.y;
              } else {
                ssdx_grid_rect rc;
                if (gt_ok && ssdx_grid_window(s_combo[q], gb, t, 0, s_combo[q].H, rc)) {
                  rx = rc.r0 | (rc.c0 << 16);
                  ry = rc.nc | (rc.nr << 16);
                }
              }
            }
            warp_for_each_cell(rx, ry, lane, [&](const int ql, const int r, const int c) {
              const ssdx_grid_combo& cb = s_combo[q0 + ql];
              const int a = cb.off + (r * cb.W + c) * cb.nb;
              const float4 ab = grid_anchor_box(cb, s_ax, r, c);
              const float iou = pair_iou(ab, anchor_area(ab), gb, ga);
              if (iou > iou_thr) {
                const int slot = s_slot[a];
                if (__float_as_uint(iou) == s_cval[slot]) atomic_min_u16(s_cg + slot, (unsigned short)g);
              }
            });
          }
        }
      }
      comp_barrier();
      PHASE_ADD(2);
      const bool overflow = *s_overflow != 0;
      if (overflow) {
        grid_fused_overflow(s_combo, nq, s_ax, s_box, s_area, s_best, s_s1a, s_slot, A, G, iou_thr);
      } else {
        // ---- closed form: only the <= G anchors that are some GT's best anchor need more than g2 ----
        for (int g = tid; g < G; g += kCompThreads) {
          const u64 p = s_best[g];
          s_s1a[g] = ((p >> 32) != 0ull) ? best_anchor(p) : -1;
        }
        comp_barrier();
        for (int g = tid; g < G; g += kCompThreads) {
          const int a = s_s1a[g];
          if (a < 0) continue;
          bool is_max = true;  // g == max(S1(a)) ?
          for (int h = g + 1; h < G; ++h) is_max = is_max && (s_s1a[h] != a);
          if (is_max) {
            int slot = s_slot[a];
            if (slot == kNone) {  // below the threshold, but the best anchor of GT box g: a new positive
              slot = atomicAdd(s_nslots, 1);  // < kSlots: pass 0 stopped at kSlots - G
              s_slot[a] = (unsigned short)slot;
              s_ccell[slot] = s_bcell[g];
              s_cg[slot] = (unsigned short)g;
            } else {
              const unsigned short g2 = s_cg[slot];
              s_cg[slot] = (s_s1a[g2] != a) ? g2 : (unsigned short)g;
            }
          }
        }
      }
      PHASE_ADD(3);
      asm volatile("bar.sync 2, %0;" ::"n"(kFThreads) : "memory");  // match done, background rows of the image in place
      PHASE_ADD(4);

      // ---- encode the positives densely and patch them over the background ----
      unsigned short* s_tab = s_slot;  // overflow path only: anchor -> matched GT
      // A positive row is 16 bytes (offsets; labels when C == 4), half of a 32-byte DRAM sector.  A 16-byte patch
      // leaves a partially written sector behind, which the memory side completes with a read-modify-write - with
      // ~360 of them per image the kernel ran 66 us per 1024 images instead of 56.  So the patch is the whole
      // sector: the positive row and its neighbour a ^ 1 (background, or positive itself - then the lower anchor of
      // the pair writes both rows), one 32-byte store when the image base is 32-byte aligned.
      const bool pair_ok = ((((size_t)b * A) & 1) == 0);  // rows a & ~1 and a | 1 share a sector
      auto matched_gt = [&](const int a) -> int {  // matched GT box of anchor a, -1 = background
        const unsigned short m = s_slot[a];
        if (m == kNone) return -1;
        return overflow ? (int)m : (int)s_cg[m];
      };
      auto emit_positive = [&](const int a, const float4 ab, const int g) {
        const float4 off_a = pair_offsets(s_box[g], anchor_centroid(ab), sx, sy, sw, sh);
        const int lab_a = s_lab[g];
        if (out_match) out_match[(size_t)b * A + a] = g;
        const int nb = a ^ 1;
        if (!pair_ok || nb >= A) {  // no sector partner: plain 16-byte rows
          st_stream_f4(offs + a, off_a);
          if (bulk_lab) {
            if (C == 4) {
              st_stream_f4(reinterpret_cast<float4*>(labs) + a, one_hot4(lab_a));
            } else {  // the background row is [1,0,...]; a label outside [0,C) gives an all-zero row (tf.one_hot)
              labs[(size_t)a * C] = (lab_a == 0) ? 1.0f : 0.0f;
              if (lab_a > 0 && lab_a < C) labs[(size_t)a * C + lab_a] = 1.0f;
            }
          }
          return;
        }
        const int g_nb = matched_gt(nb);
        if (g_nb >= 0 && nb < a) return;  // both rows positive: the lower anchor writes the sector
        float4 off_nb = make_float4(0.f, 0.f, 0.f, 0.f);
        int lab_nb = 0;
        if (g_nb >= 0) {
          off_nb = pair_offsets(s_box[g_nb], anchor_centroid(anchor_box_by_index(s_combo, nq, s_ax, nb)), sx, sy, sw,
                                sh);
          lab_nb = s_lab[g_nb];
          if (out_match) out_match[(size_t)b * A + nb] = g_nb;
        }
        const int lo = a & ~1;
        st_stream_sector(offs + lo, a == lo ? off_a : off_nb, a == lo ? off_nb : off_a);
        if (bulk_lab) {
          if (C == 4) {
            st_stream_sector(reinterpret_cast<float4*>(labs) + lo, one_hot4(a == lo ? lab_a : lab_nb),
                             one_hot4(a == lo ? lab_nb : lab_a));
          } else {
            labs[(size_t)a * C] = (lab_a == 0) ? 1.0f : 0.0f;
            if (lab_a > 0 && lab_a < C) labs[(size_t)a * C + lab_a] = 1.0f;
            if (g_nb >= 0) {
              labs[(size_t)nb * C] = (lab_nb == 0) ? 1.0f : 0.0f;
              if (lab_nb > 0 && lab_nb < C) labs[(size_t)nb * C + lab_nb] = 1.0f;
            }
          }
        }
      };
      const int n_used = min(*s_nslots, kSlots);
      if (!overflow) {
        for (int s = tid; s < n_used; s += kCompThreads) {
          const unsigned code = s_ccell[s];
          if (code == kCellDead) continue;
          const ssdx_grid_combo& cb = s_combo[code >> 24];
          const int r = (code >> 12) & 0xFFF, c = code & 0xFFF;
          emit_positive(cb.off + (r * cb.W + c) * cb.nb, grid_anchor_box(cb, s_ax, r, c), s_cg[s]);
        }
      } else {
        for (int a = tid; a < A; a += kCompThreads) {
          const unsigned short m = s_tab[a];
          if (m != kNone) emit_positive(a, anchor_box_by_index(s_combo, nq, s_ax, a), (int)m);
        }
      }
      if (!bulk_lab) {  // label rows that cannot be bulk-stored (unaligned image stride): flat coalesced walk
        const int total = A * C;
        for (int e = tid; e < total; e += kCompThreads) {
          const int a = e / C;
          const int c = e - a * C;
          int lab = 0;
          const unsigned short m = s_slot[a];  // slot, or (overflow) the matched GT itself
          if (m != kNone) lab = s_lab[overflow ? m : s_cg[m]];
          labs[e] = (lab == c) ? 1.0f : 0.0f;
        }
      }
      const int next = img + gridDim.x;
      comp_barrier();
      // ---- leave the tables empty for the next image: only what this image used ----
      if (next < B) {
        if (!overflow) {
          for (int s = tid; s < n_used; s += kCompThreads) {
            const unsigned code = s_ccell[s];
            if (code != kCellDead) {
              const ssdx_grid_combo& cb = s_combo[code >> 24];
              s_slot[cb.off + (((code >> 12) & 0xFFF) * cb.W + (code & 0xFFF)) * cb.nb] = kNone;
              s_ccell[s] = kCellDead;
            }
            s_cval[s] = 0u;
            s_cg[s] = kNone;
          }
        } else {
          for (int i = tid; i < kSlots; i += kCompThreads) {
            s_cval[i] = 0u;
            s_ccell[i] = kCellDead;
            s_cg[i] = kNone;
          }
          for (int i = tid; i < A8 / 8; i += kCompThreads) reinterpret_cast<int4*>(s_slot)[i] = make_int4(-1, -1, -1, -1);
        }
        if (tid == 0) {
          *s_nslots = 0;
          *s_overflow = 0;
        }
      }
      PHASE_ADD(5);
#ifdef SSDX_PHASE_TIMERS
      if (tid == 0) atomicAdd(&g_phase_cycles[24], 1ull);
#endif
      img = next;
    }
  }
}

// ------------------------------------------------------------------------------------------ split driver
// Small batches and large tables.  The (GT box, grid) windows of a CTA are numbered w = g * nq + q and walked by its
// warps 32 windows at a time (one per lane; the cells of the 32 windows are spread over the lanes by
// warp_for_each_cell), so the warps run independently: no block-wide scan, barriers only between the passes.
__device__ __forceinline__ unsigned fast_div_magic(const int d) { return 0xFFFFFFFFu / (unsigned)d + 1u; }
// n / d for n * d < 2^32 (window numbers: n < 2^16 * 2^8)
__device__ __forceinline__ int fast_div(const int n, const unsigned magic, const int d) {
  return d == 1 ? n : (int)__umulhi((unsigned)n, magic);
}

// K1: GT side.  grid (ceil(Gmax / kBestG), B); every (image, GT box) is owned by one CTA, which writes its packed
// best (IoU bits, ~anchor) - 0 when no anchor overlaps the box.  Dynamic shared memory: grids | axis arrays (when
// they fit; otherwise the axis arrays are read from global memory).
// Hit lists (K1 -> K2): every pair above the threshold that K1 meets is appended to the list of the span its anchor
// lies in, so that K2 does not have to search the GT boxes again.  The lists of image b live in 12 * A bytes of the
// workspace; span s owns the bytes [12 * a_lo, 12 * (a_lo + n)) of them: a 64-bit counter, then 64-bit entries
// (anchor - a_lo) << 48 | GT << 32 | IoU bits - room for ~1.5 hits per anchor of the span.  A list that overflows is
// ignored by K2, which then searches that span itself.
__host__ __device__ inline size_t span_list_offset(const int a_lo) { return ((size_t)12 * a_lo + 7) / 8 * 8; }
__host__ __device__ inline size_t span_list_stride(const int A) { return ((size_t)12 * A + 7) / 8 * 8; }  // per image
__host__ __device__ inline int span_list_capacity(const int a_lo, const int n) {
  const long long bytes = (long long)12 * (a_lo + n) - (long long)span_list_offset(a_lo) - 8;
  return bytes < 8 ? 0 : (int)(bytes / 8);
}

__global__ void ssdx_encode_grid_zero_lists_kernel(const ssdx_grid_span* __restrict__ spans, const int n_spans,
                                                   const int B, const int A, unsigned char* __restrict__ lists) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_spans * B) return;
  const int b = i / n_spans, s = i - b * n_spans;
  *reinterpret_cast<u64*>(lists + (size_t)b * span_list_stride(A) + span_list_offset(spans[s].a_lo)) = 0ull;
}

__global__ void __launch_bounds__(kThreads)
    ssdx_encode_grid_best_kernel(const ssdx_grid_combo* __restrict__ combos, const int nq,
                                 const float* __restrict__ axes, const int n_axes_smem,
                                 const ssdx_grid_span* __restrict__ spans, const int2* __restrict__ combo_span,
                                 const float4* __restrict__ gt_xyxy, const int* __restrict__ gt_count,
                                 const int A, const int Gmax, const float iou_thr, u64* __restrict__ gt_best,
                                 unsigned char* __restrict__ lists) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float4 s_box[kBestG];
  __shared__ float s_area[kBestG];
  __shared__ u64 s_best[kBestG];
  ssdx_grid_combo* s_combo = (ssdx_grid_combo*)smem_raw;
  float* s_ax = (float*)(s_combo + nq);
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);
  const int gc0 = blockIdx.x * kBestG;
  const int gn_cta = min(kBestG, G - gc0);
  if (gn_cta <= 0) return;  // unused slots of a ragged batch: K2 never reads beyond gt_count
  for (int g = tid; g < gn_cta; g += kThreads) {
    const float4 bx = __ldg(gt_xyxy + (size_t)b * Gmax + gc0 + g);
    s_box[g] = bx;
    s_area[g] = gt_area(bx);
    s_best[g] = 0ull;
  }
  for (int i = tid; i < nq * (int)(sizeof(ssdx_grid_combo) / 16); i += kThreads)
    reinterpret_cast<int4*>(s_combo)[i] = __ldg(reinterpret_cast<const int4*>(combos) + i);
  for (int i = tid; i < n_axes_smem / 4; i += kThreads)
    reinterpret_cast<int4*>(s_ax)[i] = __ldg(reinterpret_cast<const int4*>(axes) + i);
  const float* ax = n_axes_smem > 0 ? s_ax : axes;
  __syncthreads();
  const float thr_eff = fmaxf(iou_thr, 0.0f);
  const unsigned magic = fast_div_magic(nq);
  const int n_win = gn_cta * nq;
  for (int w = tid; w < n_win; w += kThreads) {
    const int g = fast_div(w, magic, nq), q = w - g * nq;
    grid_seed(s_combo[q], ax, s_box[g], s_area[g], s_best + g);
  }
  __syncthreads();
  for (int w0 = warp * 32; w0 < n_win; w0 += kThreads) {
    const int w = w0 + lane;
    int rx = 0, ry = 0, tag = 0;
    if (w < n_win) {
      const int g = fast_div(w, magic, nq), q = w - g * nq;
      tag = (g << 8) | q;
      ssdx_grid_rect rc;
      const float t = fminf(thr_eff, __uint_as_float(best_bits(s_best + g)));
      if (ssdx_grid_window(s_combo[q], s_box[g], t, 0, s_combo[q].H, rc)) {
        rx = rc.r0 | (rc.c0 << 16);
        ry = rc.nc | (rc.nr << 16);
      }
    }
    warp_for_each_cell_tagged(rx, ry, tag, lane, [&](const int tg, const int r, const int c) {
      const int g = tg >> 8;
      const ssdx_grid_combo& cb = s_combo[tg & 255];
      const float4 ab = grid_anchor_box(cb, ax, r, c);
      const float iou = pair_iou(ab, anchor_area(ab), s_box[g], s_area[g]);
      const unsigned bits = __float_as_uint(iou);
      const int a = cb.off + (r * cb.W + c) * cb.nb;
      if (iou > 0.0f && bits >= best_bits(s_best + g)) atomicMax(s_best + g, pack_best(bits, a));
      // anchor side: hand the pair to the span of the anchor.  The lanes of a round mostly hit the same one or two
      // spans: one atomicAdd per (round, span) reserves the slots of all its lanes.
      const bool hit = iou > iou_thr;
      int span_idx = -1;
      if (hit) {
        const int2 cs = __ldg(combo_span + (tg & 255));
        span_idx = cs.x + r / cs.y;
      }
      const unsigned active = __activemask();
      const unsigned peers = __match_any_sync(active, span_idx);
      if (hit) {
        const ssdx_grid_span sp = spans[span_idx];
        unsigned char* reg = lists + (size_t)b * span_list_stride(A) + span_list_offset(sp.a_lo);
        const int leader = __ffs(peers) - 1;
        const int my_rank = __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
        u64 base = 0ull;
        if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(reinterpret_cast<u64*>(reg), (u64)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        const u64 pos = base + (u64)my_rank;
        if (pos < (u64)span_list_capacity(sp.a_lo, sp.n))
          reinterpret_cast<u64*>(reg)[1 + pos] =
              ((u64)(unsigned)(a - sp.a_lo) << 48) | ((u64)(unsigned)(gc0 + g) << 32) | (u64)bits;
      }
    });
  }
  __syncthreads();
  for (int g = tid; g < gn_cta; g += kThreads) gt_best[(size_t)b * Gmax + gc0 + g] = s_best[g];
}

// K2: anchor side of one span (a row range of one feature map) of one image + the span's output rows.
// grid (n_spans, B).  Dynamic shared memory: grids of the span's map | their axis arrays | val (u32 per anchor; reused
// as the list of positives once the matches are final) | tab (u16) | s1 (u16).
__host__ __device__ inline size_t grid_span_smem_bytes(int max_span_anchors, int max_span_combos, int map_axes) {
  const size_t n8 = (size_t)(max_span_anchors + 7) / 8 * 8;
  return (size_t)max_span_combos * sizeof(ssdx_grid_combo) + (size_t)map_axes * 4 + n8 * 4 + n8 * 2 + n8 * 2 + 16;
}

__global__ void __launch_bounds__(kThreads)
    ssdx_encode_grid_span_kernel(const ssdx_grid_combo* __restrict__ combos, const ssdx_grid_span* __restrict__ spans,
                                 const float* __restrict__ axes, const int max_span_anchors,
                                 const int max_span_combos, const int map_axes_smem,
                                 const float4* __restrict__ gt_xyxy, const int* __restrict__ gt_label,
                                 const int* __restrict__ gt_count, const int A, const int Gmax, const int C,
                                 const float iou_thr, const float sx, const float sy, const float sw,
                                 const float sh, const u64* __restrict__ gt_best,
                                 const unsigned char* __restrict__ lists, float* __restrict__ out_labels,
                                 float4* __restrict__ out_offsets, int* __restrict__ out_match) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float4 s_box[kGtTile];
  __shared__ float s_area[kGtTile];
  __shared__ int s_lab[kThreads];
  __shared__ int s_npos;
  const int n8 = (max_span_anchors + 7) / 8 * 8;
  ssdx_grid_combo* s_combo = (ssdx_grid_combo*)smem_raw;
  float* s_ax = (float*)(s_combo + max_span_combos);
  unsigned* s_val = (unsigned*)(s_ax + map_axes_smem);
  unsigned short* s_tab = (unsigned short*)(s_val + n8);
  unsigned short* s_s1 = s_tab + n8;     // max(S1(a)) + 1, 0 = empty
  int* s_pos = (int*)s_val;              // compacted positives: (anchor - a_lo) << 16 | matched GT (s_val is dead by then)

  const ssdx_grid_span sp = spans[blockIdx.x];
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);
  const int nq = sp.nb;
  const u64* best = gt_best + (size_t)b * Gmax;

  for (int i = tid; i < nq * (int)(sizeof(ssdx_grid_combo) / 16); i += kThreads)
    reinterpret_cast<int4*>(s_combo)[i] = __ldg(reinterpret_cast<const int4*>(combos + sp.combo0) + i);
  for (int i = tid; i < n8 / 4; i += kThreads) reinterpret_cast<int4*>(s_val)[i] = make_int4(0, 0, 0, 0);
  for (int i = tid; i < n8 / 8; i += kThreads) {
    reinterpret_cast<int4*>(s_tab)[i] = make_int4(-1, -1, -1, -1);
    reinterpret_cast<int4*>(s_s1)[i] = make_int4(0, 0, 0, 0);
  }
  if (tid == 0) s_npos = 0;
  __syncthreads();
  // axis arrays of this map, rebased so that s_combo[q].ax indexes s_ax
  const int ax0 = s_combo[0].ax;
  const ssdx_grid_combo& lastc = s_combo[nq - 1];
  const int ax_n = lastc.ax + (2 * lastc.W + 2 * lastc.H + 3) / 4 * 4 - ax0;
  const bool ax_smem = map_axes_smem > 0 && ax_n <= map_axes_smem;
  if (ax_smem)
    for (int i = tid; i < ax_n / 4; i += kThreads)
      reinterpret_cast<int4*>(s_ax)[i] = __ldg(reinterpret_cast<const int4*>(axes + ax0) + i);
  const float* ax = ax_smem ? (s_ax - ax0) : axes;
  // S1: GT boxes whose best anchor (IoU > 0) lies in this span
  for (int g = tid; g < G; g += kThreads) {
    const u64 p = best[g];
    if ((p >> 32) != 0ull) {
      const int ba = best_anchor(p);
      if (ba >= sp.a_lo && ba < sp.a_lo + sp.n) atomic_max_u16(s_s1 + (ba - sp.a_lo), (unsigned short)(g + 1));
    }
  }

  // the pairs above the threshold, as K1 listed them for this span
  const u64* list = reinterpret_cast<const u64*>(lists + (size_t)b * span_list_stride(A) + span_list_offset(sp.a_lo));
  const u64 n_list = (G > 0) ? list[0] : 0ull;
  const bool listed = n_list <= (u64)span_list_capacity(sp.a_lo, sp.n);
  if (G > 0 && listed) {
    const int n = (int)n_list;
    for (int i = tid; i < n; i += kThreads) {  // per-anchor max IoU
      const u64 e = __ldg(list + 1 + i);
      atomicMax(s_val + (int)(e >> 48), (unsigned)e);
    }
    __syncthreads();
    for (int i = tid; i < n; i += kThreads) {  // lowest GT index among the maxima (L1 / L2 hit)
      const u64 e = __ldg(list + 1 + i);
      const int al = (int)(e >> 48);
      if ((unsigned)e == s_val[al]) atomic_min_u16(s_tab + al, (unsigned short)((e >> 32) & 0xFFFFu));
    }
  } else if (G > 0) {  // the list overflowed: search the span's pairs here
    const float thr_eff = fmaxf(iou_thr, 0.0f);
    const unsigned magic = fast_div_magic(nq);
    for (int pass = 0; pass < 2; ++pass) {  // 0: per-anchor max IoU, 1: lowest GT index among the maxima
      for (int t0 = 0; t0 < G; t0 += kGtTile) {  // GT tile staged in shared memory
        const int tn = min(kGtTile, G - t0);
        __syncthreads();
        if (pass == 0 || G > kGtTile)
          for (int g = tid; g < tn; g += kThreads) {
            const float4 bx = __ldg(gt_xyxy + (size_t)b * Gmax + t0 + g);
            s_box[g] = bx;
            s_area[g] = gt_area(bx);
          }
        __syncthreads();
        const int n_win = tn * nq;
        for (int w0 = warp * 32; w0 < n_win; w0 += kThreads) {
          const int w = w0 + lane;
          int rx = 0, ry = 0, tag = 0;
          if (w < n_win) {
            const int g = fast_div(w, magic, nq), q = w - g * nq;
            tag = (g << 8) | q;
            ssdx_grid_rect rc;  // the GT side is final (K1): only pairs above the threshold matter here
            if (ssdx_grid_window(s_combo[q], s_box[g], thr_eff, sp.row0, sp.row0 + sp.nrows, rc)) {
              rx = rc.r0 | (rc.c0 << 16);
              ry = rc.nc | (rc.nr << 16);
            }
          }
          warp_for_each_cell_tagged(rx, ry, tag, lane, [&](const int tg, const int r, const int c) {
            const int g = tg >> 8;
            const ssdx_grid_combo& cb = s_combo[tg & 255];
            const int al = cb.off + (r * cb.W + c) * cb.nb - sp.a_lo;
            const float4 ab = grid_anchor_box(cb, ax, r, c);
            const float iou = pair_iou(ab, anchor_area(ab), s_box[g], s_area[g]);
            if (iou > iou_thr) {
              const unsigned bits = __float_as_uint(iou);
              if (pass == 0) atomicMax(s_val + al, bits);
              else if (bits == s_val[al]) atomic_min_u16(s_tab + al, (unsigned short)(t0 + g));
            }
          });
        }
      }
    }
  }
  __syncthreads();

  // ---- rows of the span: background rows at once, positives compacted and encoded densely afterwards ----
  const size_t row0 = (size_t)b * A;
  for (int l0 = 0; l0 < sp.n; l0 += kThreads) {
    const int l = l0 + tid;
    int match = -1;
    if (l < sp.n) {
      const int a = sp.a_lo + l;
      const unsigned short g2 = s_tab[l];
      const int s1max = (int)s_s1[l] - 1;
      if (g2 != kNone) {  // closed form of concat -> unique -> sequential scatter
        const u64 p = best[g2];
        const bool in_s1 = ((p >> 32) != 0ull) && (best_anchor(p) == a);
        match = in_s1 ? s1max : (int)g2;
      } else {
        match = s1max;
      }
      if (match >= 0) s_pos[atomicAdd(&s_npos, 1)] = (l << 16) | match;
      else st_stream_f4(out_offsets + row0 + a, make_float4(0.f, 0.f, 0.f, 0.f));
      if (out_match) out_match[row0 + a] = match;
    }
    if (C == 4) {
      if (l < sp.n) {
        const int lab = match >= 0 ? __ldg(gt_label + (size_t)b * Gmax + match) : 0;
        st_stream_f4(reinterpret_cast<float4*>(out_labels) + row0 + sp.a_lo + l, one_hot4(lab));
      }
    } else {
      // generic class count: flat coalesced walk over the chunk's (n_here * C) floats
      __syncthreads();
      s_lab[tid] = match >= 0 ? __ldg(gt_label + (size_t)b * Gmax + match) : 0;
      __syncthreads();
      const int n_here = min(kThreads, sp.n - l0);
      float* dst = out_labels + (row0 + sp.a_lo + l0) * C;
      const int total = n_here * C;
      for (int e = tid; e < total; e += kThreads) {
        const int r = e / C;
        const int c = e - r * C;
        dst[e] = (s_lab[r] == c) ? 1.0f : 0.0f;
      }
    }
  }
  __syncthreads();
  const int npos = s_npos;
  const int Wm = s_combo[0].W;
  for (int i = tid; i < npos; i += kThreads) {
    const int l = s_pos[i] >> 16, pm = s_pos[i] & 0xFFFF;
    const int cell = l / nq, k = l - cell * nq;          // the span starts at a row boundary of its map
    const int r = sp.row0 + cell / Wm, c = cell % Wm;
    const float4 ab = grid_anchor_box(s_combo[k], ax, r, c);
    st_stream_f4(out_offsets + row0 + sp.a_lo + l,
                 pair_offsets(__ldg(gt_xyxy + (size_t)b * Gmax + pm), anchor_centroid(ab), sx, sy, sw, sh));
  }
}

constexpr int kAxesSmemMax = 12288;  // floats of axis arrays a CTA stages in shared memory (48 KB)

}  // namespace

// ---- host side (called by ssdx_encode) ----
bool ssdx_encode_grid_applies(const ssdx_anchors* anchors, int gmax, int num_classes, float iou_threshold) {
  // a negative threshold makes every anchor a positive (handled by the tile drivers); NaN compares false everywhere
  return anchors->n_combos > 0 && iou_threshold >= 0.0f && gmax >= 0 && gmax <= 32767 && num_classes >= 1;
}

bool ssdx_encode_grid_fused_fits(const ssdx_anchors* anchors, int gmax, int num_classes) {
  return gmax <= kFusedMaxG && anchors->A <= kGridFusedMaxA && anchors->n_axes <= kAxesSmemMax / 2 &&
         grid_fused_smem_bytes(anchors->A, gmax, num_classes, anchors->n_combos, anchors->n_axes) <= 110 * 1024;
}

int ssdx_encode_grid_launch(const ssdx_anchors* anchors, const bool fused, const float* gt_xyxy,
                            const int32_t* gt_label, const int32_t* gt_count, int32_t batch, int32_t gmax,
                            int32_t num_classes, float iou_threshold, const float* std_host, float* out_labels,
                            float* out_offsets, int32_t* out_match, void* workspace, cudaStream_t s) {
  const int A = anchors->A;
  const int nq = anchors->n_combos;
  if (fused) {
    const size_t smem = grid_fused_smem_bytes(A, gmax, num_classes, nq, anchors->n_axes);
    SSDX_CUDA(cudaFuncSetAttribute(ssdx_encode_grid_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    static thread_local size_t cached_smem = 0;  // resident CTAs per SM for this shared-memory size
    static thread_local int cached_per_sm = 0;
    if (cached_smem != smem || cached_per_sm == 0) {
      SSDX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cached_per_sm, ssdx_encode_grid_fused_kernel, kFThreads,
                                                              smem));
      cached_smem = smem;
    }
    const int per_sm = cached_per_sm;
    const int grid = std::min((int)batch, std::max(1, per_sm) * ssdx_sm_count(anchors->device));
    ssdx_encode_grid_fused_kernel<<<grid, kFThreads, smem, s>>>(
        anchors->combos, nq, anchors->axes, anchors->n_axes, (const float4*)gt_xyxy, gt_label, gt_count, batch, A,
        gmax, num_classes, iou_threshold, std_host[0], std_host[1], std_host[2], std_host[3], out_labels,
        (float4*)out_offsets, out_match);
    SSDX_LAUNCH_CHECK("ssdx_encode_grid_fused_kernel");
    return SSDX_OK;
  }
  const int gm = gmax > 0 ? gmax : 1;
  u64* gt_best = (u64*)workspace;
  unsigned char* lists = (unsigned char*)workspace + ssdx_align_up((size_t)batch * gm * sizeof(u64), 256);
  if (gmax > 0) {
    const int nz = anchors->n_spans * batch;
    ssdx_encode_grid_zero_lists_kernel<<<(nz + 255) / 256, 256, 0, s>>>(anchors->spans, anchors->n_spans, batch, A,
                                                                        lists);
    SSDX_LAUNCH_CHECK("ssdx_encode_grid_zero_lists_kernel");
    const int ax1 = anchors->n_axes <= kAxesSmemMax ? anchors->n_axes : 0;
    const size_t smem1 = (size_t)nq * sizeof(ssdx_grid_combo) + (size_t)ax1 * 4;
    SSDX_CUDA(cudaFuncSetAttribute(ssdx_encode_grid_best_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem1));
    dim3 g1((gm + kBestG - 1) / kBestG, batch);
    ssdx_encode_grid_best_kernel<<<g1, kThreads, smem1, s>>>(anchors->combos, nq, anchors->axes, ax1, anchors->spans,
                                                             anchors->combo_span, (const float4*)gt_xyxy, gt_count,
                                                             A, gm, iou_threshold, gt_best, lists);
    SSDX_LAUNCH_CHECK("ssdx_encode_grid_best_kernel");
  }
  const int ax2 = anchors->max_map_axes <= kAxesSmemMax ? anchors->max_map_axes : 0;
  const size_t smem2 = grid_span_smem_bytes(anchors->max_span_anchors, anchors->max_span_combos, ax2);
  SSDX_CUDA(cudaFuncSetAttribute(ssdx_encode_grid_span_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem2));
  dim3 g2(anchors->n_spans, batch);
  ssdx_encode_grid_span_kernel<<<g2, kThreads, smem2, s>>>(
      anchors->combos, anchors->spans, anchors->axes, anchors->max_span_anchors, anchors->max_span_combos, ax2,
      (const float4*)gt_xyxy, gt_label, gt_count, A, gm, num_classes, iou_threshold, std_host[0], std_host[1],
      std_host[2], std_host[3], gt_best, lists, out_labels, (float4*)out_offsets, out_match);
  SSDX_LAUNCH_CHECK("ssdx_encode_grid_span_kernel");
  return SSDX_OK;
}
