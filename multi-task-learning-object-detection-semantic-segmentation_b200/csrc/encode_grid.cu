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
// Drivers:
//   fused : one CTA per image (batch >= number of SMs, A small enough for per-anchor tables in shared memory).  The
//           constant background rows of the whole image leave as TMA bulk stores at the start (as in encode.cu);
//           seeds -> windows -> block scan -> cells (pass 0: per-anchor max IoU and per-GT best anchor; pass 1:
//           lowest GT index among the maxima) -> closed form -> the positives are patched over the background.
//   split : K1 grid (GT chunks, images) computes the per-GT best anchors, K2 grid (row spans of a feature map,
//           images) the anchor side of its span and writes the span's rows: small batches and large tables
//           (dense stress: 100k anchors x 256 GT boxes).
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
#define SSDX_GRID_MINB 2
#endif
#ifndef SSDX_GRID_BGROWS
#define SSDX_GRID_BGROWS 512
#endif
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kWinCap = 512;                 // windows per round (GT boxes of the round x grids)
constexpr int kBgRows = SSDX_GRID_BGROWS;    // background rows per bulk store
constexpr int kPosCap = 1024;                // compacted positives per image (more are encoded in place)
constexpr int kBestG = 32;                   // GT boxes per CTA of the split driver's K1
constexpr int kGtTile = 256;                 // GT boxes staged per round by the split driver's K2
constexpr unsigned short kNone = 0xFFFFu;
constexpr int kGridFusedMaxA = 12288;        // per-anchor tables of the fused driver: 6 bytes per anchor

// datacoder.py:210-220 for one pair; 0 when the boxes do not intersect (the reference's max(0, .) makes inter 0)
__device__ __forceinline__ float pair_iou(const float4 ab, const float aa, const float4 gb, const float ga) {
  const float iw = __fadd_rn(__fsub_rn(fminf(ab.z, gb.z), fmaxf(ab.x, gb.x)), 1.0f);
  const float ih = __fadd_rn(__fsub_rn(fminf(ab.w, gb.w), fmaxf(ab.y, gb.y)), 1.0f);
  if (!(iw > 0.0f && ih > 0.0f)) return 0.0f;
  const float inter = __fmul_rn(iw, ih);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ga), inter));
}

// seed of the GT-side bound: exact IoU of the anchor of this grid nearest to the GT centre
__device__ __forceinline__ void grid_seed(const ssdx_grid_combo& cb, const float4* __restrict__ corners,
                                          const float* __restrict__ area, const float4 gb, const float ga,
                                          u64* s_best_g) {
  const float gw = gb.z - gb.x + 1.0f, gh = gb.w - gb.y + 1.0f;
  if (!(gw > 0.0f && gh > 0.0f)) return;
  const float fc = rintf((0.5f * (gb.x + gb.z) - cb.cx0) * cb.inv_dx);
  const float fr = rintf((0.5f * (gb.y + gb.w) - cb.cy0) * cb.inv_dy);
  const int c = (int)fminf(fmaxf(fc, 0.0f), (float)(cb.W - 1));
  const int r = (int)fminf(fmaxf(fr, 0.0f), (float)(cb.H - 1));
  const int a = cb.off + (r * cb.W + c) * cb.nb;
  const float iou = pair_iou(__ldg(corners + a), __ldg(area + a), gb, ga);
  if (iou > 0.0f) atomicMax(s_best_g, pack_best(__float_as_uint(iou), a));
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

// Exclusive prefix sum of s[0..n) in place, n <= kWinCap; s[n] = total (also returned).  Called by all threads, the
// counts must be complete (barrier) before the call; ends with a barrier.
__device__ __forceinline__ int block_scan_excl(int* s, const int n, int* s_wsum) {
  constexpr int P = kWinCap / kThreads;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int base = tid * P;
  int v[P];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < P; ++k) {
    v[k] = (base + k < n) ? s[base + k] : 0;
    sum += v[k];
  }
  int incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int u = __shfl_up_sync(kFull, incl, d);
    if (lane >= d) incl += u;
  }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = lane < kWarps ? s_wsum[lane] : 0;
    int wi = w;
#pragma unroll
    for (int d = 1; d < kWarps; d <<= 1) {
      const int u = __shfl_up_sync(kFull, wi, d);
      if (lane >= d) wi += u;
    }
    if (lane < kWarps) s_wsum[lane] = wi - w;
    if (lane == kWarps - 1) s_wsum[kWarps] = wi;
  }
  __syncthreads();
  int run = s_wsum[warp] + incl - sum;
#pragma unroll
  for (int k = 0; k < P; ++k) {
    if (base + k < n) s[base + k] = run;
    run += v[k];
  }
  const int total = s_wsum[kWarps];
  if (tid == 0) s[n] = total;
  __syncthreads();
  return total;
}

// largest w in [0, n) with s_start[w] <= i  (s_start is non-decreasing, s_start[0] = 0 <= i < s_start[n])
__device__ __forceinline__ int find_window(const int* s_start, const int n, const int i) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (s_start[mid] <= i) lo = mid;
    else hi = mid;
  }
  return lo;
}

// Windows of one round: GT boxes [0, gn) of the staged tile x grids [0, nq).  Window w = g * nq + q.
//   T_MODE 0: t(g) = min(thr_eff, current GT-side best)   (fused driver, K1)
//   T_MODE 1: t(g) = thr_eff                              (K2: the GT side is already final)
template <int T_MODE>
__device__ __forceinline__ void make_windows(const ssdx_grid_combo* s_combo, const int nq, const float4* s_box,
                                             const u64* s_best, const int g_first, const int gn,
                                             const float thr_eff, const int row_lo_rel, const int nrows_clip,
                                             int* s_cnt, int2* s_rect) {
  const int n_win = gn * nq;
  for (int w = threadIdx.x; w < n_win; w += kThreads) {
    const int gl = w / nq, q = w - gl * nq;
    const int g = g_first + gl;
    const ssdx_grid_combo& cb = s_combo[q];
    float t = thr_eff;
    if (T_MODE == 0) t = fminf(t, __uint_as_float(best_bits(s_best + g)));
    ssdx_grid_rect r;
    const int row_lo = nrows_clip > 0 ? row_lo_rel : 0;
    const int row_hi = nrows_clip > 0 ? row_lo_rel + nrows_clip : cb.H;
    int cnt = 0;
    if (ssdx_grid_window(cb, s_box[g], t, row_lo, row_hi, r)) {
      cnt = r.nr * r.nc;
      s_rect[w] = make_int2(r.r0 | (r.c0 << 16), r.nc);
    }
    s_cnt[w] = cnt;
  }
}

// ------------------------------------------------------------------------------------------ fused driver
__host__ __device__ inline int grid_label_rows(int C) {  // rows per constant label block (multiple of 4)
  int r = ((kBgRows * 16) / (4 * C)) & ~3;
  return r < 4 ? 4 : r;
}
__host__ __device__ inline bool grid_bulk_labels(int A, int C) { return ((A * (long long)C) % 4) == 0 && C <= 1024; }

__host__ __device__ inline size_t grid_fused_smem_bytes(int A, int G, int C, int nq) {
  const size_t g = (size_t)(G > 0 ? G : 1);
  size_t n = (size_t)kBgRows * 16;                                                  // bg_off
  n += grid_bulk_labels(A, C) ? (size_t)grid_label_rows(C) * C * 4 : 16;            // bg_lab
  n += g * 16 + g * 8 + g * 4 + g * 4 + g * 4;                                      // box, best, area, lab, s1a
  n = (n + 15) / 16 * 16;
  n += (size_t)nq * sizeof(ssdx_grid_combo);                                        // grids
  n += (size_t)kWinCap * 8 + (size_t)(kWinCap + 4) * 4;                             // rect, cnt/start
  n += (size_t)kPosCap * 4 + 64;                                                    // pos, ctrl + warp sums
  n += (size_t)((A + 3) / 4 * 4) * 4;                                               // val
  n += (size_t)((A + 7) / 8 * 8) * 2;                                               // tab
  return n + 16;
}

__global__ void __launch_bounds__(kThreads, SSDX_GRID_MINB)
    ssdx_encode_grid_fused_kernel(const float4* __restrict__ corners, const float* __restrict__ area,
                                  const float4* __restrict__ anc_cent, const ssdx_grid_combo* __restrict__ combos,
                                  const int nq, const float4* __restrict__ gt_xyxy,
                                  const int* __restrict__ gt_label, const int* __restrict__ gt_count, const int A,
                                  const int Gmax, const int C, const float iou_thr, const float sx, const float sy,
                                  const float sw, const float sh, float* __restrict__ out_labels,
                                  float4* __restrict__ out_offsets, int* __restrict__ out_match) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);
  const int Gs = Gmax > 0 ? Gmax : 1;
  const bool bulk_lab = grid_bulk_labels(A, C);
  const int lab_rows = grid_label_rows(C);
  const int A4 = (A + 3) / 4 * 4, A8 = (A + 7) / 8 * 8;

  float4* bg_off;
  float* bg_lab;
  float4* s_box;
  u64* s_best;
  float* s_area;
  int* s_lab;
  int* s_s1a;
  ssdx_grid_combo* s_combo;
  int2* s_rect;
  int* s_cnt;
  int* s_pos;
  int* s_ctrl;
  unsigned* s_val;
  unsigned short* s_tab;
  {
    unsigned char* p = smem_raw;
    bg_off = (float4*)p;  p += (size_t)kBgRows * 16;
    bg_lab = (float*)p;   p += bulk_lab ? (size_t)lab_rows * C * 4 : 16;
    s_box = (float4*)p;   p += (size_t)Gs * 16;
    s_best = (u64*)p;     p += (size_t)Gs * 8;
    s_area = (float*)p;   p += (size_t)Gs * 4;
    s_lab = (int*)p;      p += (size_t)Gs * 4;
    s_s1a = (int*)p;      p += (size_t)Gs * 4;
    p = smem_raw + ((size_t)(p - smem_raw) + 15) / 16 * 16;
    s_combo = (ssdx_grid_combo*)p; p += (size_t)nq * sizeof(ssdx_grid_combo);
    s_rect = (int2*)p;    p += (size_t)kWinCap * 8;
    s_cnt = (int*)p;      p += (size_t)(kWinCap + 4) * 4;
    s_pos = (int*)p;      p += (size_t)kPosCap * 4;
    s_ctrl = (int*)p;     p += 64;
    s_val = (unsigned*)p; p += (size_t)A4 * 4;
    s_tab = (unsigned short*)p;
  }
  int* s_wsum = s_ctrl + 4;  // kWarps + 1 ints

  // ---- phase 0: constant background blocks, GT staging, tables ----
  PHASE_T0();
  for (int i = tid; i < kBgRows; i += kThreads) bg_off[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bulk_lab)
    for (int r = tid; r < lab_rows; r += kThreads) {
      bg_lab[r * C] = 1.0f;
      for (int c = 1; c < C; ++c) bg_lab[r * C + c] = 0.0f;
    }
  for (int g = tid; g < G; g += kThreads) {
    const float4 bx = __ldg(gt_xyxy + (size_t)b * Gmax + g);
    s_box[g] = bx;
    s_area[g] = gt_area(bx);
    s_lab[g] = __ldg(gt_label + (size_t)b * Gmax + g);
    s_best[g] = 0ull;
  }
  for (int i = tid; i < nq * (int)(sizeof(ssdx_grid_combo) / 16); i += kThreads)
    reinterpret_cast<int4*>(s_combo)[i] = __ldg(reinterpret_cast<const int4*>(combos) + i);
  for (int i = tid; i < A4 / 4; i += kThreads) reinterpret_cast<int4*>(s_val)[i] = make_int4(0, 0, 0, 0);
  for (int i = tid; i < A8 / 8; i += kThreads) reinterpret_cast<int4*>(s_tab)[i] = make_int4(-1, -1, -1, -1);
  if (tid == 0) s_ctrl[1] = 0;  // number of positives
  fence_async_smem();           // the background blocks are read by the async proxy
  __syncthreads();
  PHASE_ADD(0);

  float4* offs = out_offsets + (size_t)b * A;
  float* labs = out_labels + (size_t)b * A * C;
  if (warp == kWarps - 1 && lane == 0) {  // background for the whole image, drained by the TMA unit during the match
    const int n_off = (A + kBgRows - 1) / kBgRows;
    const int n_lab = bulk_lab ? (A + lab_rows - 1) / lab_rows : 0;
    for (int i = 0; i < n_off + n_lab; ++i) {
      if (i < n_off) {
        const int r0 = i * kBgRows;
        bulk_store(offs + r0, bg_off, (unsigned)(min(kBgRows, A - r0) * 16));
      } else {
        const int r0 = (i - n_off) * lab_rows;
        bulk_store(labs + (size_t)r0 * C, bg_lab, (unsigned)(min(lab_rows, A - r0) * C * 4));
      }
    }
    bulk_commit();
  }
  __syncwarp();

  if (G > 0) {
    const float thr_eff = fmaxf(iou_thr, 0.0f);
    // ---- seeds: nearest anchor of every grid ----
    for (int w = tid; w < G * nq; w += kThreads) {
      const int g = w / nq, q = w - g * nq;
      grid_seed(s_combo[q], corners, area, s_box[g], s_area[g], s_best + g);
    }
    __syncthreads();
    PHASE_ADD(1);
    // ---- windows -> scan -> cells; pass 0: per-anchor max IoU + GT side, pass 1: lowest GT among the maxima ----
    const int g_round = max(1, kWinCap / nq);  // GT boxes per round
    const bool one_round = G <= g_round;
    for (int pass = 0; pass < 2; ++pass) {
      for (int g0 = 0; g0 < G; g0 += g_round) {
        const int gn = min(g_round, G - g0);
        const int n_win = gn * nq;
        int total;
        if (pass == 0 || !one_round) {
          make_windows<0>(s_combo, nq, s_box, s_best, g0, gn, thr_eff, 0, 0, s_cnt, s_rect);
          __syncthreads();
          total = block_scan_excl(s_cnt, n_win, s_wsum);
        } else {
          total = s_cnt[n_win];  // the windows of pass 0 are still in place
        }
        for (int i = tid; i < total; i += kThreads) {
          const int w = find_window(s_cnt, n_win, i);
          const int gl = w / nq, q = w - gl * nq;
          const int g = g0 + gl;
          const ssdx_grid_combo& cb = s_combo[q];
          const int2 rc = s_rect[w];
          const int local = i - s_cnt[w];
          const int rr = local / rc.y, cc = local - rr * rc.y;
          const int a = cb.off + (((rc.x & 0xFFFF) + rr) * cb.W + (rc.x >> 16) + cc) * cb.nb;
          const float iou = pair_iou(__ldg(corners + a), __ldg(area + a), s_box[g], s_area[g]);
          const unsigned bits = __float_as_uint(iou);
          if (pass == 0) {
            if (iou > iou_thr) atomicMax(s_val + a, bits);
            if (iou > 0.0f && bits >= best_bits(s_best + g)) atomicMax(s_best + g, pack_best(bits, a));
          } else {
            if (iou > iou_thr && bits == s_val[a]) atomic_min_u16(s_tab + a, (unsigned short)g);
          }
        }
        __syncthreads();
      }
    }
  }
  PHASE_ADD(2);
  if (warp == kWarps - 1 && lane == 0) bulk_wait_all();  // background rows are in place before any patch
  __syncthreads();
  PHASE_ADD(3);

  // ---- closed form: only the <= G anchors that are some GT's best anchor need more than g2 ----
  for (int g = tid; g < G; g += kThreads) {
    const u64 p = s_best[g];
    s_s1a[g] = ((p >> 32) != 0ull) ? best_anchor(p) : -1;
  }
  __syncthreads();
  for (int g = tid; g < G; g += kThreads) {
    const int a = s_s1a[g];
    if (a < 0) continue;
    bool is_max = true;  // g == max(S1(a)) ?
    for (int h = g + 1; h < G; ++h) is_max = is_max && (s_s1a[h] != a);
    if (is_max) {
      const unsigned short g2 = s_tab[a];
      s_tab[a] = (g2 != kNone && s_s1a[g2] != a) ? g2 : (unsigned short)g;
    }
  }
  __syncthreads();
  PHASE_ADD(4);

  // ---- compact the positives, encode them densely, patch them over the background ----
  auto emit_positive = [&](const int a, const int g) {
    st_stream_f4(offs + a, pair_offsets(s_box[g], __ldg(anc_cent + a), sx, sy, sw, sh));
    if (bulk_lab) {
      const int lab = s_lab[g];
      if (C == 4) {
        const float4 oh = make_float4(lab == 0 ? 1.f : 0.f, lab == 1 ? 1.f : 0.f, lab == 2 ? 1.f : 0.f,
                                      lab == 3 ? 1.f : 0.f);
        st_stream_f4(reinterpret_cast<float4*>(labs) + a, oh);
      } else {  // the background row is [1,0,...]; a label outside [0,C) gives an all-zero row (tf.one_hot)
        labs[(size_t)a * C] = (lab == 0) ? 1.0f : 0.0f;
        if (lab > 0 && lab < C) labs[(size_t)a * C + lab] = 1.0f;
      }
    }
  };
  for (int i = tid; i < A8 / 8; i += kThreads) {
    const int4 v = reinterpret_cast<const int4*>(s_tab)[i];
    if ((v.x & v.y & v.z & v.w) == -1) continue;  // 8 background rows
    const int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const unsigned m = (k & 1) ? ((unsigned)w[k >> 1] >> 16) : ((unsigned)w[k >> 1] & 0xFFFFu);
      if (m != kNone) {
        const int a = i * 8 + k;
        const int slot = atomicAdd(&s_ctrl[1], 1);
        if (slot < kPosCap) s_pos[slot] = (a << 8) | (int)m;
        else emit_positive(a, (int)m);
      }
    }
  }
  __syncthreads();
  const int npos = min(s_ctrl[1], kPosCap);
  for (int s = tid; s < npos; s += kThreads) emit_positive(s_pos[s] >> 8, s_pos[s] & 255);

  if (!bulk_lab) {  // label rows that cannot be bulk-stored (unaligned image stride): flat coalesced walk
    const int total = A * C;
    for (int e = tid; e < total; e += kThreads) {
      const int a = e / C;
      const int c = e - a * C;
      const unsigned short m = s_tab[a];
      labs[e] = ((m != kNone ? s_lab[m] : 0) == c) ? 1.0f : 0.0f;
    }
  }
  if (out_match)
    for (int a = tid; a < A; a += kThreads) {
      const unsigned short m = s_tab[a];
      out_match[(size_t)b * A + a] = (m != kNone) ? (int)m : -1;
    }
  PHASE_ADD(5);
#ifdef SSDX_PHASE_TIMERS
  if (tid == 0) atomicAdd(&g_phase_cycles[24], 1ull);
#endif
}

// ------------------------------------------------------------------------------------------ split driver
// K1: GT side.  grid (ceil(Gmax / kBestG), B); every (image, GT box) is owned by one CTA, which writes its packed
// best (IoU bits, ~anchor) - 0 when no anchor overlaps the box.
__global__ void __launch_bounds__(kThreads)
    ssdx_encode_grid_best_kernel(const float4* __restrict__ corners, const float* __restrict__ area,
                                 const ssdx_grid_combo* __restrict__ combos, const int nq,
                                 const float4* __restrict__ gt_xyxy, const int* __restrict__ gt_count,
                                 const int Gmax, const float iou_thr, u64* __restrict__ gt_best) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float4 s_box[kBestG];
  __shared__ float s_area[kBestG];
  __shared__ u64 s_best[kBestG];
  __shared__ int2 s_rect[kWinCap];
  __shared__ int s_cnt[kWinCap + 4];
  __shared__ int s_wsum[kWarps + 1];
  ssdx_grid_combo* s_combo = (ssdx_grid_combo*)smem_raw;
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);
  const int gc0 = blockIdx.x * kBestG;
  const int gn_cta = min(kBestG, G - gc0);
  if (gn_cta <= 0) {  // unused slots of a ragged batch: K2 never reads beyond gt_count, nothing to write
    return;
  }
  for (int g = tid; g < gn_cta; g += kThreads) {
    const float4 bx = __ldg(gt_xyxy + (size_t)b * Gmax + gc0 + g);
    s_box[g] = bx;
    s_area[g] = gt_area(bx);
    s_best[g] = 0ull;
  }
  for (int i = tid; i < nq * (int)(sizeof(ssdx_grid_combo) / 16); i += kThreads)
    reinterpret_cast<int4*>(s_combo)[i] = __ldg(reinterpret_cast<const int4*>(combos) + i);
  __syncthreads();
  const float thr_eff = fmaxf(iou_thr, 0.0f);
  for (int w = tid; w < gn_cta * nq; w += kThreads) {
    const int g = w / nq, q = w - g * nq;
    grid_seed(s_combo[q], corners, area, s_box[g], s_area[g], s_best + g);
  }
  __syncthreads();
  const int g_round = max(1, kWinCap / nq);
  for (int g0 = 0; g0 < gn_cta; g0 += g_round) {
    const int gn = min(g_round, gn_cta - g0);
    const int n_win = gn * nq;
    make_windows<0>(s_combo, nq, s_box, s_best, g0, gn, thr_eff, 0, 0, s_cnt, s_rect);
    __syncthreads();
    const int total = block_scan_excl(s_cnt, n_win, s_wsum);
    for (int i = tid; i < total; i += kThreads) {
      const int w = find_window(s_cnt, n_win, i);
      const int gl = w / nq, q = w - gl * nq;
      const int g = g0 + gl;
      const ssdx_grid_combo& cb = s_combo[q];
      const int2 rc = s_rect[w];
      const int local = i - s_cnt[w];
      const int rr = local / rc.y, cc = local - rr * rc.y;
      const int a = cb.off + (((rc.x & 0xFFFF) + rr) * cb.W + (rc.x >> 16) + cc) * cb.nb;
      const float iou = pair_iou(__ldg(corners + a), __ldg(area + a), s_box[g], s_area[g]);
      const unsigned bits = __float_as_uint(iou);
      if (iou > 0.0f && bits >= best_bits(s_best + g)) atomicMax(s_best + g, pack_best(bits, a));
    }
    __syncthreads();
  }
  for (int g = tid; g < gn_cta; g += kThreads) gt_best[(size_t)b * Gmax + gc0 + g] = s_best[g];
}

// K2: anchor side of one span (a row range of one feature map) of one image + the span's output rows.
// grid (n_spans, B).  Dynamic shared memory: combos of the span's map | val (u32 per anchor) | tab (u16) | s1 (u16).
__host__ __device__ inline size_t grid_span_smem_bytes(int max_span_anchors, int max_span_combos) {
  const size_t n4 = (size_t)(max_span_anchors + 7) / 8 * 8;
  return (size_t)max_span_combos * sizeof(ssdx_grid_combo) + n4 * 4 + n4 * 2 + n4 * 2 + n4 * 4 + 16;
}

__global__ void __launch_bounds__(kThreads)
    ssdx_encode_grid_span_kernel(const float4* __restrict__ corners, const float* __restrict__ area,
                                 const float4* __restrict__ anc_cent, const ssdx_grid_combo* __restrict__ combos,
                                 const ssdx_grid_span* __restrict__ spans, const int max_span_anchors,
                                 const int max_span_combos, const float4* __restrict__ gt_xyxy,
                                 const int* __restrict__ gt_label, const int* __restrict__ gt_count, const int A,
                                 const int Gmax, const int C, const float iou_thr, const float sx, const float sy,
                                 const float sw, const float sh, const u64* __restrict__ gt_best,
                                 float* __restrict__ out_labels, float4* __restrict__ out_offsets,
                                 int* __restrict__ out_match) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float4 s_box[kGtTile];
  __shared__ float s_area[kGtTile];
  __shared__ int2 s_rect[kWinCap];
  __shared__ int s_cnt[kWinCap + 4];
  __shared__ int s_wsum[kWarps + 1];
  __shared__ int s_npos;
  const int n4 = (max_span_anchors + 7) / 8 * 8;
  ssdx_grid_combo* s_combo = (ssdx_grid_combo*)smem_raw;
  unsigned* s_val = (unsigned*)(s_combo + max_span_combos);
  unsigned short* s_tab = (unsigned short*)(s_val + n4);
  unsigned short* s_s1 = s_tab + n4;     // max(S1(a)) + 1, 0 = empty
  int* s_pos = (int*)(s_s1 + n4);        // compacted positives: (anchor - a_lo) << 16 | matched GT

  const ssdx_grid_span sp = spans[blockIdx.x];
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);
  const int nq = sp.nb;
  const u64* best = gt_best + (size_t)b * Gmax;

  for (int i = tid; i < nq * (int)(sizeof(ssdx_grid_combo) / 16); i += kThreads)
    reinterpret_cast<int4*>(s_combo)[i] = __ldg(reinterpret_cast<const int4*>(combos + sp.combo0) + i);
  for (int i = tid; i < n4 / 4; i += kThreads) reinterpret_cast<int4*>(s_val)[i] = make_int4(0, 0, 0, 0);
  for (int i = tid; i < n4 / 8; i += kThreads) {
    reinterpret_cast<int4*>(s_tab)[i] = make_int4(-1, -1, -1, -1);
    reinterpret_cast<int4*>(s_s1)[i] = make_int4(0, 0, 0, 0);
  }
  if (tid == 0) s_npos = 0;
  __syncthreads();
  // S1: GT boxes whose best anchor (IoU > 0) lies in this span
  for (int g = tid; g < G; g += kThreads) {
    const u64 p = best[g];
    if ((p >> 32) != 0ull) {
      const int ba = best_anchor(p);
      if (ba >= sp.a_lo && ba < sp.a_lo + sp.n) atomic_max_u16(s_s1 + (ba - sp.a_lo), (unsigned short)(g + 1));
    }
  }

  if (G > 0) {
    const float thr_eff = fmaxf(iou_thr, 0.0f);
    const int g_round = max(1, kWinCap / nq);
    for (int pass = 0; pass < 2; ++pass) {
      for (int t0 = 0; t0 < G; t0 += kGtTile) {  // GT tile staged in shared memory
        const int tn = min(kGtTile, G - t0);
        __syncthreads();
        for (int g = tid; g < tn; g += kThreads) {
          const float4 bx = __ldg(gt_xyxy + (size_t)b * Gmax + t0 + g);
          s_box[g] = bx;
          s_area[g] = gt_area(bx);
        }
        __syncthreads();
        for (int g0 = 0; g0 < tn; g0 += g_round) {
          const int gn = min(g_round, tn - g0);
          const int n_win = gn * nq;
          make_windows<1>(s_combo, nq, s_box, nullptr, g0, gn, thr_eff, sp.row0, sp.nrows, s_cnt, s_rect);
          __syncthreads();
          const int total = block_scan_excl(s_cnt, n_win, s_wsum);
          for (int i = tid; i < total; i += kThreads) {
            const int w = find_window(s_cnt, n_win, i);
            const int gl = w / nq, q = w - gl * nq;
            const int g = g0 + gl;
            const ssdx_grid_combo& cb = s_combo[q];
            const int2 rc = s_rect[w];
            const int local = i - s_cnt[w];
            const int rr = local / rc.y, cc = local - rr * rc.y;
            const int a = cb.off + (((rc.x & 0xFFFF) + rr) * cb.W + (rc.x >> 16) + cc) * cb.nb;
            const float iou = pair_iou(__ldg(corners + a), __ldg(area + a), s_box[g], s_area[g]);
            if (iou > iou_thr) {
              const unsigned bits = __float_as_uint(iou);
              if (pass == 0) atomicMax(s_val + (a - sp.a_lo), bits);
              else if (bits == s_val[a - sp.a_lo]) atomic_min_u16(s_tab + (a - sp.a_lo), (unsigned short)(t0 + g));
            }
          }
          __syncthreads();
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
        const float4 oh = make_float4(lab == 0 ? 1.f : 0.f, lab == 1 ? 1.f : 0.f, lab == 2 ? 1.f : 0.f,
                                      lab == 3 ? 1.f : 0.f);
        st_stream_f4(reinterpret_cast<float4*>(out_labels) + row0 + sp.a_lo + l, oh);
      }
    } else {
      // generic class count: flat coalesced walk over the chunk's (n_here * C) floats
      __shared__ int s_lab[kThreads];
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
  for (int i = tid; i < npos; i += kThreads) {
    const int pa = sp.a_lo + (s_pos[i] >> 16), pm = s_pos[i] & 0xFFFF;
    st_stream_f4(out_offsets + row0 + pa,
                 pair_offsets(__ldg(gt_xyxy + (size_t)b * Gmax + pm), __ldg(anc_cent + pa), sx, sy, sw, sh));
  }
}

}  // namespace

// ---- host side (called by ssdx_encode) ----
bool ssdx_encode_grid_applies(const ssdx_anchors* anchors, int gmax, int num_classes, float iou_threshold) {
  // a negative threshold makes every anchor a positive (handled by the tile drivers); NaN compares false everywhere
  return anchors->n_combos > 0 && iou_threshold >= 0.0f && gmax >= 0 && gmax <= 32767 && num_classes >= 1;
}

bool ssdx_encode_grid_fused_fits(const ssdx_anchors* anchors, int gmax, int num_classes) {
  return gmax <= kFusedMaxG && anchors->A <= kGridFusedMaxA &&
         grid_fused_smem_bytes(anchors->A, gmax, num_classes, anchors->n_combos) <= 110 * 1024;
}

size_t ssdx_encode_grid_workspace_bytes(int batch, int gmax) {
  return ssdx_align_up((size_t)batch * (size_t)(gmax > 0 ? gmax : 1) * sizeof(u64), 256);
}

int ssdx_encode_grid_launch(const ssdx_anchors* anchors, const bool fused, const float* gt_xyxy,
                            const int32_t* gt_label, const int32_t* gt_count, int32_t batch, int32_t gmax,
                            int32_t num_classes, float iou_threshold, const float* std_host, float* out_labels,
                            float* out_offsets, int32_t* out_match, void* workspace, cudaStream_t s) {
  const int A = anchors->A;
  const int nq = anchors->n_combos;
  if (fused) {
    const size_t smem = grid_fused_smem_bytes(A, gmax, num_classes, nq);
    SSDX_CUDA(cudaFuncSetAttribute(ssdx_encode_grid_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    ssdx_encode_grid_fused_kernel<<<batch, kThreads, smem, s>>>(
        anchors->corners, anchors->area, anchors->enc_cent, anchors->combos, nq, (const float4*)gt_xyxy, gt_label,
        gt_count, A, gmax, num_classes, iou_threshold, std_host[0], std_host[1], std_host[2], std_host[3],
        out_labels, (float4*)out_offsets, out_match);
    SSDX_LAUNCH_CHECK("ssdx_encode_grid_fused_kernel");
    return SSDX_OK;
  }
  const int gm = gmax > 0 ? gmax : 1;
  u64* gt_best = (u64*)workspace;
  if (gmax > 0) {
    dim3 g1((gm + kBestG - 1) / kBestG, batch);
    ssdx_encode_grid_best_kernel<<<g1, kThreads, (size_t)nq * sizeof(ssdx_grid_combo), s>>>(
        anchors->corners, anchors->area, anchors->combos, nq, (const float4*)gt_xyxy, gt_count, gm, iou_threshold,
        gt_best);
    SSDX_LAUNCH_CHECK("ssdx_encode_grid_best_kernel");
  }
  const size_t smem2 = grid_span_smem_bytes(anchors->max_span_anchors, anchors->max_span_combos);
  SSDX_CUDA(cudaFuncSetAttribute(ssdx_encode_grid_span_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem2));
  dim3 g2(anchors->n_spans, batch);
  ssdx_encode_grid_span_kernel<<<g2, kThreads, smem2, s>>>(
      anchors->corners, anchors->area, anchors->enc_cent, anchors->combos, anchors->spans,
      anchors->max_span_anchors, anchors->max_span_combos, (const float4*)gt_xyxy, gt_label, gt_count, A, gm,
      num_classes, iou_threshold, std_host[0], std_host[1], std_host[2], std_host[3], gt_best, out_labels,
      (float4*)out_offsets, out_match);
  SSDX_LAUNCH_CHECK("ssdx_encode_grid_span_kernel");
  return SSDX_OK;
}
