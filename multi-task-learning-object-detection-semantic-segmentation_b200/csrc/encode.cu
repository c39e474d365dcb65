// SSD target encode (reference: datacoder.py:205-300).
//
// Parity contract: IoU is evaluated with exactly the reference's fp32 expression tree
// (one rounding per op, IEEE division, association (area_a + area_g) - inter); both arg-max
// reductions resolve ties to the lowest index; the concat -> UniqueV2 -> sequential scatter of
// :244/:286 is replaced by its closed form:
//     S1(a) = { g : argmax_a iou[:,g] == a  and  max_a iou[:,g] > 0 }
//     g2(a) = argmax_g iou[a,:]  if  max_g iou[a,:] > thr
//     match(a) = g2 if g2 exists and g2 not in S1(a);  else max(S1(a)) if non-empty;  else background.
//
// The A x G IoU matrix is never materialised and most of it is never evaluated.  Anchors are walked in the
// "match order" built by ssdx_anchors_create (sorted by area half-octave, then Morton code of the centre; tiles of
// kSsdxTile anchors with a bounding box and an area range).  A pair can only matter if
//     iou > thr                 (anchor side, datacoder.py:236-241), or
//     iou >= running best of g  (GT side, :230-231; '>=' because ties resolve to the lower anchor index),
// i.e. if iou >= t(g) = min(max(thr,0), best(g)).  Three conservative tests (each one only ever skips pairs that
// provably fail it; fp32 rounding is monotone, slack factors cover the rounding of the bounds themselves):
//   (i)   tile bounding box misses the GT box           -> inter == 0 for every anchor of the tile;
//   (ii)  area ratio: iou <= min(area)/max(area)         -> tile skipped if even that is < t(g);
//   (iii) pair: iou <= inter / max(area_a, area_g)       -> division skipped if inter <= t(g)*max(area_g, tile min area).
// (i)+(ii) are evaluated by 32 lanes for 32 GT boxes at once (ballot -> surviving GT list); (iii) costs one
// compare per pair.  IoU == 0 pairs can never win a strict '>' against a running best that starts at (0, index 0).
//
// Two drivers share the pair loop:
//   fused : one CTA per image (reference configuration).  98 % of the output rows are the constant background
//           row, so the CTA starts by issuing TMA bulk stores (cp.async.bulk shared -> global) of constant
//           background blocks for the whole image; they drain to HBM while the warps match.  After the match the
//           few positive rows (~180 of 9600) are compacted, encoded densely and patched over the background.
//   tiled : grid (anchor tiles, images) + a second emit kernel, for large A (dense stress: 100k anchors);
//           the per-GT arg-max crosses CTAs through one 64-bit atomicMax per (CTA, GT).
#include <stdlib.h>

#include "common.cuh"
#include "encode_common.cuh"

#ifdef SSDX_PHASE_TIMERS
static __device__ unsigned long long g_phase_cycles[32];
extern "C" __attribute__((visibility("default"))) int ssdx_debug_phase_cycles_encode(unsigned long long* out32) {
  unsigned long long zero[32] = {0};
  if (cudaMemcpyFromSymbol(out32, g_phase_cycles, sizeof(zero)) != cudaSuccess) return -2;
  if (cudaMemcpyToSymbol(g_phase_cycles, zero, sizeof(zero)) != cudaSuccess) return -2;
  return 0;
}
#endif

namespace {

using namespace ssdx_enc;

constexpr int kApt = kSsdxApt;      // anchors per lane per tile
#ifndef SSDX_ENC_THREADS
#define SSDX_ENC_THREADS 256
#endif
#ifndef SSDX_ENC_MINB
#define SSDX_ENC_MINB (1024 / SSDX_ENC_THREADS)  // resident CTAs per SM the register budget is sized for
#endif
#ifndef SSDX_ENC_BGROWS
#define SSDX_ENC_BGROWS 512
#endif
constexpr int kFusedThreads = SSDX_ENC_THREADS;  // one CTA per image
constexpr int kTiledThreads = 256;  // grid (tile groups, images)
constexpr int kTiledTilesPerCta = 16;
constexpr int kGtTile = 256;        // GT boxes staged per shared-memory tile (tiled driver)
constexpr int kEmitThreads = 256;
constexpr int kEmitSpan = 8;  // chunks of kEmitThreads anchors per emit CTA
constexpr int kEmitAnchors = kEmitSpan * kEmitThreads;
constexpr int kPosCap = 1024;   // compacted positives per image (more are encoded in place)
constexpr int kBgRows = SSDX_ENC_BGROWS;  // background rows per bulk store (offsets: 16 B per row)
constexpr float kSlackUp = 1.0001f, kSlackDown = 0.9999f;

struct AnchorTile {
  float4 box[kApt];
  float area[kApt];
  int idx[kApt];  // original anchor index, -1 = padding slot (empty box)
  float best_v[kApt];
  int best_g[kApt];
};

// lane l owns the match-order slots tile*kSsdxTile + u*32 + l
__device__ __forceinline__ void load_tile(AnchorTile& t, const float4* __restrict__ m_box,
                                          const float* __restrict__ m_area, const int* __restrict__ m_idx,
                                          const int tile, const int lane) {
#pragma unroll
  for (int u = 0; u < kApt; ++u) {
    const int i = tile * kSsdxTile + u * 32 + lane;
    t.box[u] = __ldg(m_box + i);
    t.area[u] = __ldg(m_area + i);
    t.idx[u] = __ldg(m_idx + i);
    t.best_v[u] = 0.0f;  // np.argmax over an all-zero row -> index 0
    t.best_g[u] = 0;
  }
}

// One anchor tile against gn GT boxes in shared memory (s_best: CTA-shared packed running best per GT).
// (shared-memory pointers: no __restrict__, see nms.cu)
__device__ __forceinline__ void match_tile(AnchorTile& t, const float4 tb, const float2 ta, const float4* s_box,
                                           const float* s_area, u64* s_best, const int gn, const int g0,
                                           const float thr_eff, const int lane) {
  for (int c0 = 0; c0 < gn; c0 += 32) {
    // ---- tests (i) and (ii): lane l decides for GT box c0 + l ----
    const int gl = c0 + lane;
    bool keep = false;
    float tgp_lane = 0.0f;
    if (gl < gn) {
      const float4 gb = s_box[gl];
      const float ga = s_area[gl];
      const float tw = __fadd_rn(__fsub_rn(fminf(tb.z, gb.z), fmaxf(tb.x, gb.x)), 1.0f);
      const float th = __fadd_rn(__fsub_rn(fminf(tb.w, gb.w), fmaxf(tb.y, gb.y)), 1.0f);
      const float tg = fminf(thr_eff, __uint_as_float(best_bits(s_best + gl)));  // >= 0
      // test (ii): iou <= min(amax, ga, inter(tile box, gt)) / max(amin, ga), written without a division
      const float gm = fmaxf(ga, ta.x);
      const float num = fminf(fminf(ta.y, ga), __fmul_rn(tw, th));
      keep = (tw > 0.0f) && (th > 0.0f) && (__fmul_rn(num, kSlackUp) >= __fmul_rn(tg, gm));
      tgp_lane = __fmul_rn(__fmul_rn(tg, gm), kSlackDown);
    }
    unsigned mask = __ballot_sync(kFull, keep);
#ifdef SSDX_PHASE_TIMERS
    if (lane == 0) atomicAdd(&g_phase_cycles[25], (unsigned long long)__popc(mask));  // surviving (tile, GT)
#endif
    while (mask != 0u) {  // ascending GT index: the anchor-side arg-max keeps the lower GT on ties
      const int src = __ffs(mask) - 1;
      mask &= mask - 1u;
      const int g = c0 + src;
      const float4 gb = s_box[g];
      const float tgp = __shfl_sync(kFull, tgp_lane, src);
      float inter[kApt];
      bool cand[kApt];
      bool any = false;
#pragma unroll
      for (int u = 0; u < kApt; ++u) {
        // datacoder.py:210-216 without the max(0, .): a pair with iw <= 0 or ih <= 0 is never a candidate
        const float iw = __fadd_rn(__fsub_rn(fminf(t.box[u].z, gb.z), fmaxf(t.box[u].x, gb.x)), 1.0f);
        const float ih = __fadd_rn(__fsub_rn(fminf(t.box[u].w, gb.w), fmaxf(t.box[u].y, gb.y)), 1.0f);
        inter[u] = __fmul_rn(iw, ih);
        cand[u] = (iw > 0.0f) && (inter[u] > tgp);  // tgp >= 0: implies ih > 0.  Test (iii).
        any = any || cand[u];
      }
      if (!__any_sync(kFull, any)) continue;
#ifdef SSDX_PHASE_TIMERS
      if (lane == 0) atomicAdd(&g_phase_cycles[26], 1ull);  // slow-path entries
#endif
      const float ga = s_area[g];
      const unsigned cur = best_bits(s_best + g);
      const float tg = fminf(thr_eff, __uint_as_float(cur));
      const float tg1 = __fadd_rn(1.0f, tg), tgs = __fmul_rn(tg, kSlackDown);
      float mv = 0.0f;
      int ma = 0x7FFFFFFF;
#pragma unroll
      for (int u = 0; u < kApt; ++u) {
        // iou >= tg  <=>  inter * (1 + tg) >= tg * (area_a + area_g): the same decision without the division,
        // up to rounding (slack on the right-hand side), taken with the current bound
        const float asum = __fadd_rn(t.area[u], ga);
        if (cand[u] && __fmul_rn(inter[u], tg1) >= __fmul_rn(asum, tgs)) {
          // datacoder.py:220  inter / ((area_a + area_g) - inter), IEEE division
          const float iou = __fdiv_rn(inter[u], __fsub_rn(asum, inter[u]));
#ifdef SSDX_PHASE_TIMERS
          atomicAdd(&g_phase_cycles[27], 1ull);  // divisions
#endif
          if (iou > t.best_v[u]) {  // strict: ties keep the lower GT index
            t.best_v[u] = iou;
            t.best_g[u] = g0 + g;
          }
          if (iou > mv || (iou == mv && t.idx[u] < ma)) {  // ties -> lower original anchor index
            mv = iou;
            ma = t.idx[u];
          }
        }
      }
      // GT side: only when some lane reaches the running best (IoU >= 0, so the bit pattern is monotone)
      const unsigned mvb = __float_as_uint(mv);
      if (!__any_sync(kFull, mvb != 0u && mvb >= cur)) continue;
#ifdef SSDX_PHASE_TIMERS
      if (lane == 0) atomicAdd(&g_phase_cycles[28], 1ull);  // GT-side updates
#endif
      const unsigned mbits = __reduce_max_sync(kFull, mvb);
      const int amin = __reduce_min_sync(kFull, (mvb == mbits) ? ma : 0x7FFFFFFF);
      if (lane == 0) atomicMax(s_best + g, pack_best(mbits, amin));
    }
  }
}

// Seed of the GT-side bound: one warp evaluates, for GT box g, the tile that looks most promising (largest
// bound (ii), then nearest centre) and publishes the best real pair it finds.  Any real pair is a valid lower bound
// of max_a iou[a,g]; a good one lets tests (ii) and (iii) cull from the first tile on.
// (t_bbox / t_area: shared memory in the fused driver, global in the tiled one: generic loads, no __restrict__)
__device__ __forceinline__ void seed_gt_best(const float4* __restrict__ m_box, const float* __restrict__ m_area,
                                             const int* __restrict__ m_idx, const float4* t_bbox,
                                             const float2* t_area, const int n_tiles, const float4 gb,
                                             const float ga, u64* s_best_g, const int lane) {
  float best_score = -1.0f;
  int best_tile = 0;
  const float gcx = 0.5f * (gb.x + gb.z), gcy = 0.5f * (gb.y + gb.w);
  for (int tile = lane; tile < n_tiles; tile += 32) {
    const float4 tb = t_bbox[tile];
    const float2 ta = t_area[tile];
    const float tw = fminf(tb.z, gb.z) - fmaxf(tb.x, gb.x) + 1.0f;
    const float th = fminf(tb.w, gb.w) - fmaxf(tb.y, gb.y) + 1.0f;
    if (tw > 0.0f && th > 0.0f) {
      const float ub = fminf(fminf(ta.y, ga), tw * th) / fmaxf(fmaxf(ta.x, ga), 1e-30f);
      const float dx = 0.5f * (tb.x + tb.z) - gcx, dy = 0.5f * (tb.y + tb.w) - gcy;
      const float score = ub / (1.0f + (dx * dx + dy * dy) / fmaxf(ga, 1e-30f));  // heuristic only
      if (score > best_score) {
        best_score = score;
        best_tile = tile;
      }
    }
  }
  const unsigned sb = __reduce_max_sync(kFull, __float_as_uint(fmaxf(best_score, 0.0f)));
  if (sb == 0u) return;  // no tile touches this GT box
  const int tile = __reduce_min_sync(kFull, (__float_as_uint(fmaxf(best_score, 0.0f)) == sb) ? best_tile : 0x7FFFFFFF);
  float mv = 0.0f;
  int ma = 0x7FFFFFFF;
#pragma unroll
  for (int u = 0; u < kApt; ++u) {
    const int i = tile * kSsdxTile + u * 32 + lane;
    const float4 ab = __ldg(m_box + i);
    const float iw = __fadd_rn(__fsub_rn(fminf(ab.z, gb.z), fmaxf(ab.x, gb.x)), 1.0f);
    const float ih = __fadd_rn(__fsub_rn(fminf(ab.w, gb.w), fmaxf(ab.y, gb.y)), 1.0f);
    const float inter = __fmul_rn(iw, ih);
    if (iw > 0.0f && inter > 0.0f) {
      const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(__ldg(m_area + i), ga), inter));
      const int idx = __ldg(m_idx + i);
      if (iou > mv || (iou == mv && idx < ma)) {
        mv = iou;
        ma = idx;
      }
    }
  }
  const unsigned mvb = __float_as_uint(mv);
  const unsigned mbits = __reduce_max_sync(kFull, mvb);
  if (mbits == 0u) return;
  const int amin = __reduce_min_sync(kFull, (mvb == mbits) ? ma : 0x7FFFFFFF);
  if (lane == 0) atomicMax(s_best_g, pack_best(mbits, amin));
}

// ------------------------------------------------------------------------------------------ fused driver
__host__ __device__ inline int fused_label_rows(int C) {  // rows per constant label block (multiple of 4)
  int r = ((kBgRows * 16) / (4 * C)) & ~3;
  return r < 4 ? 4 : r;
}
__host__ __device__ inline bool fused_bulk_labels(int A, int C) { return ((A * (long long)C) % 4) == 0 && C <= 1024; }

__host__ __device__ inline size_t fused_smem_bytes(int A, int G, int C) {
  const size_t g = (size_t)(G > 0 ? G : 1);
  size_t n = (size_t)kBgRows * 16;                                                   // bg_off
  n += fused_bulk_labels(A, C) ? (size_t)fused_label_rows(C) * C * 4 : 16;           // bg_lab
  n += g * 16 + g * 8 + g * 4 + g * 4 + g * 4;                                       // box, best, area, lab, s1a
  n = (n + 15) / 16 * 16;
  n += (size_t)kPosCap * 4 + 16;                                                     // pos, ctrl
  n += (size_t)((A + kSsdxTile - 1) / kSsdxTile) * 24 + 16;                          // tile boxes + area ranges
  n += (size_t)((A + 7) / 8 * 8) * 2;                                                // table
  return n + 16;
}

__global__ void __launch_bounds__(kFusedThreads, SSDX_ENC_MINB)
    ssdx_encode_fused_kernel(const float4* __restrict__ m_box, const float* __restrict__ m_area,
                             const int* __restrict__ m_idx, const float4* __restrict__ t_bbox,
                             const float2* __restrict__ t_area, const int n_tiles,
                             const float4* __restrict__ anc_cent, const float4* __restrict__ gt_xyxy,
                             const int* __restrict__ gt_label, const int* __restrict__ gt_count, const int A,
                             const int Gmax, const int C, const float iou_thr, const float sx, const float sy,
                             const float sw, const float sh, float* __restrict__ out_labels,
                             float4* __restrict__ out_offsets, int* __restrict__ out_match) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kFusedThreads / 32;
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);
  const int Gs = Gmax > 0 ? Gmax : 1;
  const bool bulk_lab = fused_bulk_labels(A, C);
  const int lab_rows = fused_label_rows(C);
  const int A8 = (A + 7) / 8 * 8;

  float4* bg_off;
  float* bg_lab;
  float4* s_box;
  u64* s_best;
  float* s_area;
  int* s_lab;
  int* s_s1a;
  int* s_pos;
  int* s_ctrl;
  float4* s_tbox;
  float2* s_tarea;
  short* s_tab;
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
    s_pos = (int*)p;      p += (size_t)kPosCap * 4;
    s_ctrl = (int*)p;     p += 16;
    s_tbox = (float4*)p;  p += (size_t)n_tiles * 16;
    s_tarea = (float2*)p; p += ((size_t)n_tiles * 8 + 15) / 16 * 16;
    s_tab = (short*)p;
  }

  // ---- phase 0: constant background blocks, GT staging, tables ----
  PHASE_T0();
  for (int i = tid; i < kBgRows; i += kFusedThreads) bg_off[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bulk_lab)
    for (int r = tid; r < lab_rows; r += kFusedThreads) {
      bg_lab[r * C] = 1.0f;
      for (int c = 1; c < C; ++c) bg_lab[r * C + c] = 0.0f;
    }
  for (int g = tid; g < G; g += kFusedThreads) {
    const float4 bx = __ldg(gt_xyxy + (size_t)b * Gmax + g);
    s_box[g] = bx;
    s_area[g] = gt_area(bx);
    s_lab[g] = __ldg(gt_label + (size_t)b * Gmax + g);
    s_best[g] = 0ull;
  }
  for (int i = tid; i < A8 / 8; i += kFusedThreads) reinterpret_cast<int4*>(s_tab)[i] = make_int4(-1, -1, -1, -1);
  for (int t = tid; t < n_tiles; t += kFusedThreads) {  // tile table: read by every seed warp and every tile visit
    s_tbox[t] = __ldg(t_bbox + t);
    s_tarea[t] = __ldg(t_area + t);
  }
  if (tid == 0) {
    s_ctrl[0] = kWarps;  // next tile to hand out (the first kWarps tiles are taken statically)
    s_ctrl[1] = 0;       // number of positives
  }
  fence_async_smem();  // the background blocks are read by the async proxy
  __syncthreads();
  PHASE_ADD(16);  // init

  float4* offs = out_offsets + (size_t)b * A;
  float* labs = out_labels + (size_t)b * A * C;
#ifndef SSDX_EXP_SKIP_BULK
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
#endif
  __syncwarp();

  // ---- phase 1: match (dynamic tile hand-out; tile 0 holds the largest anchors) ----
#ifdef SSDX_EXP_SKIP_MATCH
  if (G > 1000000) {
#else
  if (G > 0) {
#endif
    const float thr_eff = fmaxf(iou_thr, 0.0f);
#ifndef SSDX_EXP_SKIP_SEED
    for (int g = warp; g < G; g += kWarps)
      seed_gt_best(m_box, m_area, m_idx, s_tbox, s_tarea, n_tiles, s_box[g], s_area[g], s_best + g, lane);
#endif
    PHASE_ADD(17);  // seed (no barrier: the other warps' seeds arrive through s_best whenever they are ready)
    int tile = warp;
    while (tile < n_tiles) {
      AnchorTile t;
      load_tile(t, m_box, m_area, m_idx, tile, lane);
      match_tile(t, s_tbox[tile], s_tarea[tile], s_box, s_area, s_best, G, 0, thr_eff, lane);
#pragma unroll
      for (int u = 0; u < kApt; ++u)
        if (t.idx[u] >= 0 && t.best_v[u] > iou_thr) s_tab[t.idx[u]] = (short)t.best_g[u];
      int nxt = 0;
      if (lane == 0) nxt = atomicAdd(&s_ctrl[0], 1);
      tile = __shfl_sync(kFull, nxt, 0);
    }
  }
  PHASE_ADD(18);  // match (warp 0)
  if (warp == kWarps - 1 && lane == 0) bulk_wait_all();  // background rows are in place before any patch
  __syncthreads();
  PHASE_ADD(19);  // wait for the other warps and the bulk stores

  // ---- closed form: only the <= G anchors that are some GT's best anchor need more than g2 ----
  for (int g = tid; g < G; g += kFusedThreads) {
    const u64 p = s_best[g];
    s_s1a[g] = ((p >> 32) != 0ull) ? best_anchor(p) : -1;
  }
  __syncthreads();
  for (int g = tid; g < G; g += kFusedThreads) {
    const int a = s_s1a[g];
    if (a < 0) continue;
    bool is_max = true;  // g == max(S1(a)) ?
    for (int h = g + 1; h < G; ++h) is_max = is_max && (s_s1a[h] != a);
    if (is_max) {
      const int g2 = s_tab[a];
      s_tab[a] = (short)((g2 >= 0 && s_s1a[g2] != a) ? g2 : g);
    }
  }
  __syncthreads();

  PHASE_ADD(20);  // closed form
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
  for (int i = tid; i < A8 / 8; i += kFusedThreads) {
    const int4 v = reinterpret_cast<const int4*>(s_tab)[i];
    if (((unsigned)(v.x & v.y & v.z & v.w) & 0x80008000u) == 0x80008000u) continue;  // 8 background rows
    const int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int m = (int)(short)((k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xFFFF));
      if (m >= 0) {
        const int a = i * 8 + k;
        const int slot = atomicAdd(&s_ctrl[1], 1);
        if (slot < kPosCap) s_pos[slot] = (a << 8) | m;
        else emit_positive(a, m);
      }
    }
  }
  __syncthreads();
  const int npos = min(s_ctrl[1], kPosCap);
  for (int s = tid; s < npos; s += kFusedThreads) emit_positive(s_pos[s] >> 8, s_pos[s] & 255);

  if (!bulk_lab) {  // label rows that cannot be bulk-stored (unaligned image stride): flat coalesced walk
    const int total = A * C;
    for (int e = tid; e < total; e += kFusedThreads) {
      const int a = e / C;
      const int c = e - a * C;
      const int m = s_tab[a];
      labs[e] = ((m >= 0 ? s_lab[m] : 0) == c) ? 1.0f : 0.0f;
    }
  }
  if (out_match)
    for (int a = tid; a < A; a += kFusedThreads) out_match[(size_t)b * A + a] = s_tab[a];
  PHASE_ADD(21);  // compact + patch
#ifdef SSDX_PHASE_TIMERS
  if (tid == 0) atomicAdd(&g_phase_cycles[24], 1ull);
#endif
}

// ------------------------------------------------------------------------------------------ tiled driver
// closed form of concat -> unique -> sequential scatter for one anchor
__device__ __forceinline__ int resolve_match(int a, int g2, int s1max, const u64* gbest) {
  if (g2 >= 0) {
    const u64 p = gbest[g2];
    const bool in_s1 = ((p >> 32) != 0ull) && (best_anchor(p) == a);
    return in_s1 ? s1max : g2;
  }
  return s1max;
}

// Seeds of the GT-side bound for the tiled driver: one warp per (image, GT box), grid (ceil(Gmax / warps), B).
__global__ void __launch_bounds__(kTiledThreads)
    ssdx_encode_seed_kernel(const float4* __restrict__ m_box, const float* __restrict__ m_area,
                            const int* __restrict__ m_idx, const float4* __restrict__ t_bbox,
                            const float2* __restrict__ t_area, const int n_tiles,
                            const float4* __restrict__ gt_xyxy, const int* __restrict__ gt_count, const int Gmax,
                            u64* __restrict__ gt_best) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (kTiledThreads / 32) + (threadIdx.x >> 5);
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);
  if (g >= G) return;
  const float4 gb = __ldg(gt_xyxy + (size_t)b * Gmax + g);
  seed_gt_best(m_box, m_area, m_idx, t_bbox, t_area, n_tiles, gb, gt_area(gb), gt_best + (size_t)b * Gmax + g, lane);
}

// grid (ceil(n_tiles / kTiledTilesPerCta), B).  anchor_g2 is pre-set to -1; only anchors above the threshold write.
__global__ void __launch_bounds__(kTiledThreads)
    ssdx_encode_match_kernel(const float4* __restrict__ m_box, const float* __restrict__ m_area,
                             const int* __restrict__ m_idx, const float4* __restrict__ t_bbox,
                             const float2* __restrict__ t_area, const int n_tiles,
                             const float4* __restrict__ gt_xyxy, const int* __restrict__ gt_count, const int A,
                             const int Gmax, const float iou_thr, u64* __restrict__ gt_best,
                             int* __restrict__ anchor_g2) {
  __shared__ float4 s_box[kGtTile];
  __shared__ float s_area[kGtTile];
  __shared__ u64 s_best[kGtTile];
  __shared__ u64 s_seen[kGtTile];
  __shared__ float s_bv[kTiledTilesPerCta * kSsdxTile];  // per-anchor running best across GT tiles
  __shared__ int s_bg[kTiledTilesPerCta * kSsdxTile];

  constexpr int kWarps = kTiledThreads / 32;
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);
  if (G == 0) return;
  const int tile0 = blockIdx.x * kTiledTilesPerCta;
  const int tiles_here = min(kTiledTilesPerCta, n_tiles - tile0);
  const float thr_eff = fmaxf(iou_thr, 0.0f);
  const bool multi = G > kGtTile;

  for (int g0 = 0; g0 < G; g0 += kGtTile) {
    const int gn = min(kGtTile, G - g0);
    __syncthreads();
    for (int g = tid; g < gn; g += kTiledThreads) {
      const float4 bx = __ldg(gt_xyxy + (size_t)b * Gmax + g0 + g);
      s_box[g] = bx;
      s_area[g] = gt_area(bx);
      // whatever other CTAs have published so far is a true IoU of this GT box: a valid lower bound
      const u64 seen = *((volatile u64*)(gt_best + (size_t)b * Gmax + g0 + g));
      s_best[g] = seen;
      s_seen[g] = seen;
    }
    __syncthreads();
    for (int k = warp; k < tiles_here; k += kWarps) {
      AnchorTile t;
      load_tile(t, m_box, m_area, m_idx, tile0 + k, lane);
      if (multi && g0 > 0) {
#pragma unroll
        for (int u = 0; u < kApt; ++u) {
          t.best_v[u] = s_bv[k * kSsdxTile + u * 32 + lane];
          t.best_g[u] = s_bg[k * kSsdxTile + u * 32 + lane];
        }
      }
      match_tile(t, __ldg(t_bbox + tile0 + k), __ldg(t_area + tile0 + k), s_box, s_area, s_best, gn, g0, thr_eff,
                 lane);
      if (g0 + kGtTile >= G) {  // last GT tile: publish the anchors above the threshold
#pragma unroll
        for (int u = 0; u < kApt; ++u)
          if (t.idx[u] >= 0 && t.best_v[u] > iou_thr) anchor_g2[(size_t)b * A + t.idx[u]] = t.best_g[u];
      } else {
#pragma unroll
        for (int u = 0; u < kApt; ++u) {
          s_bv[k * kSsdxTile + u * 32 + lane] = t.best_v[u];
          s_bg[k * kSsdxTile + u * 32 + lane] = t.best_g[u];
        }
      }
    }
    __syncthreads();
    // one global atomic per (CTA, GT) that improved on what it saw
    for (int g = tid; g < gn; g += kTiledThreads)
      if (s_best[g] > s_seen[g]) atomicMax(gt_best + (size_t)b * Gmax + g0 + g, s_best[g]);
  }
}

__global__ void __launch_bounds__(kEmitThreads)
    ssdx_encode_emit_kernel(const float4* __restrict__ anc_cent, const float4* __restrict__ gt_xyxy,
                            const int* __restrict__ gt_label, const int* __restrict__ gt_count, const int A,
                            const int Gmax, const int C, const float sx, const float sy, const float sw,
                            const float sh, const u64* __restrict__ gt_best, const int* __restrict__ anchor_g2,
                            float* __restrict__ out_labels, float4* __restrict__ out_offsets,
                            int* __restrict__ out_match) {
  // one CTA per span of kEmitSpan * kEmitThreads consecutive anchors of one image: the scan over the image's GT boxes
  // (step-1 winners that fall into the span) and the barriers are paid once per span
  __shared__ int s_s1max[kEmitAnchors];
  __shared__ int s_pos[kEmitAnchors];  // compacted positives of the span: (anchor - a_base) << 16 | matched GT
  __shared__ int s_lab[kEmitThreads];
  __shared__ int s_npos;
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int a_base = blockIdx.x * kEmitAnchors;
  int G = gt_count[b];
  G = G < 0 ? 0 : (G > Gmax ? Gmax : G);
  const u64* best = gt_best + (size_t)b * Gmax;

#pragma unroll
  for (int k = 0; k < kEmitSpan; ++k) s_s1max[k * kEmitThreads + tid] = -1;
  if (tid == 0) s_npos = 0;
  __syncthreads();
  // S1: GT boxes whose best anchor (IoU > 0) lies in this span
  for (int g = tid; g < G; g += kEmitThreads) {
    const u64 p = best[g];
    if ((p >> 32) != 0ull) {
      const int ba = best_anchor(p);
      if (ba >= a_base && ba < a_base + kEmitAnchors) atomicMax(&s_s1max[ba - a_base], g);
    }
  }
  __syncthreads();

  const size_t row0 = (size_t)b * A;
  int g2[kEmitSpan];
#pragma unroll
  for (int k = 0; k < kEmitSpan; ++k) {  // kEmitSpan independent loads in flight
    const int a = a_base + k * kEmitThreads + tid;
    g2[k] = (a < A) ? __ldcs(anchor_g2 + row0 + a) : -1;
  }
#pragma unroll
  for (int k = 0; k < kEmitSpan; ++k) {
    const int al = k * kEmitThreads + tid;
    const int a = a_base + al;
    int match = -1;
    if (a < A) match = resolve_match(a, g2[k], s_s1max[al], best);
    int lab = 0;  // background row = one_hot(0)
    if (match >= 0) lab = __ldg(gt_label + (size_t)b * Gmax + match);
    // offsets: background rows are written at once; the positives (log + six IEEE divisions each) are compacted and
    // encoded densely afterwards instead of diverging every warp of the CTA
    if (a < A) {
      if (match >= 0) s_pos[atomicAdd(&s_npos, 1)] = (al << 16) | match;  // al < 2^15, match <= 32766
      else st_stream_f4(out_offsets + row0 + a, make_float4(0.f, 0.f, 0.f, 0.f));
      if (out_match) out_match[row0 + a] = match;
    }
    if (C == 4) {
      if (a < A) {
        const float4 oh = make_float4(lab == 0 ? 1.f : 0.f, lab == 1 ? 1.f : 0.f, lab == 2 ? 1.f : 0.f,
                                      lab == 3 ? 1.f : 0.f);
        st_stream_f4(reinterpret_cast<float4*>(out_labels) + row0 + a, oh);
      }
    } else {
      // generic class count: flat coalesced walk over the chunk's (n_here * C) floats
      __syncthreads();
      s_lab[tid] = lab;
      __syncthreads();
      const int c_base = a_base + k * kEmitThreads;
      const int n_here = min(kEmitThreads, A - c_base);
      if (n_here > 0) {
        float* dst = out_labels + (row0 + c_base) * C;
        const int total = n_here * C;
        for (int e = tid; e < total; e += kEmitThreads) {
          const int l = e / C;
          const int c = e - l * C;
          dst[e] = (s_lab[l] == c) ? 1.0f : 0.0f;
        }
      }
    }
  }
  __syncthreads();
  const int npos = s_npos;
  for (int i = tid; i < npos; i += kEmitThreads) {
    const int pa = a_base + (s_pos[i] >> 16), pm = s_pos[i] & 0xFFFF;
    st_stream_f4(out_offsets + row0 + pa,
                 pair_offsets(__ldg(gt_xyxy + (size_t)b * Gmax + pm), __ldg(anc_cent + pa), sx, sy, sw, sh));
  }
}

inline bool use_fused(int A, int gmax, int C, int batch, int sms, int opt) {
  // one CTA per image needs enough images to fill the GPU, and A/G small enough for shared memory
  const bool fits = gmax <= kFusedMaxG && A <= 32768 && fused_smem_bytes(A, gmax, C) <= 100 * 1024;
  if (opt == SSDX_ENC_TILE_TILED) return false;
  if (opt == SSDX_ENC_TILE_FUSED) return fits;
  return fits && batch >= sms;
}

}  // namespace

extern "C" size_t ssdx_encode_workspace_bytes(int32_t num_anchors, int32_t batch, int32_t gmax) {
  if (num_anchors <= 0 || batch <= 0 || gmax < 0) return 0;
  const size_t best = ssdx_align_up((size_t)batch * (size_t)(gmax > 0 ? gmax : 1) * sizeof(u64), 256);
  // the larger of: matched GT per anchor (tile drivers, 4 bytes per anchor) and the hit lists of the split grid driver
  // (12 bytes per anchor, encode_grid.cu)
  const size_t per_anchor = ssdx_align_up((size_t)batch * (((size_t)num_anchors * 12 + 7) / 8 * 8), 256);
  return best + per_anchor;
}

extern "C" int ssdx_encode(const ssdx_anchors* anchors, const float* gt_xyxy, const int32_t* gt_label,
                           const int32_t* gt_count, int32_t batch, int32_t gmax, int32_t num_classes,
                           float iou_threshold, const float* std_host, float* out_labels, float* out_offsets,
                           int32_t* out_match, void* workspace, size_t workspace_bytes, ssdx_stream stream) {
  SSDX_CHECK_ARG(anchors != nullptr, "ssdx_encode: anchors is NULL");
  SSDX_CHECK_ARG(batch >= 0 && gmax >= 0, "ssdx_encode: batch and gmax must be >= 0");
  SSDX_CHECK_ARG(num_classes >= 1, "ssdx_encode: num_classes must be >= 1");
  SSDX_CHECK_ARG(std_host != nullptr && out_labels != nullptr && out_offsets != nullptr && gt_count != nullptr,
                 "ssdx_encode: NULL pointer argument");
  SSDX_CHECK_ARG(gmax == 0 || (gt_xyxy != nullptr && gt_label != nullptr), "ssdx_encode: GT arrays are NULL");
  SSDX_CHECK_ARG(((uintptr_t)gt_xyxy & 15) == 0 && ((uintptr_t)out_offsets & 15) == 0 &&
                     ((uintptr_t)out_labels & 15) == 0,
                 "ssdx_encode: gt_xyxy/out_labels/out_offsets must be 16-byte aligned");
  SSDX_CHECK_ARG(batch <= 65535, "ssdx_encode: batch must be <= 65535 per call (got %d)", batch);
  SSDX_CHECK_ARG(gmax <= 32767, "ssdx_encode: at most 32767 GT boxes per image (got %d)", gmax);
  if (batch == 0) return SSDX_OK;
  const int A = anchors->A;
  const size_t need = ssdx_encode_workspace_bytes(A, batch, gmax);
  if (workspace == nullptr || workspace_bytes < need) {
    ssdx_set_error("ssdx_encode: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return SSDX_ERR_WORKSPACE;
  }
  SSDX_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "ssdx_encode: workspace must be 256-byte aligned");
  SSDX_DEVICE(anchors->device);
  cudaStream_t s = (cudaStream_t)stream;
  const int gm = gmax > 0 ? gmax : 1;

  const int opt = ssdx_option_encode_path();  // ssdx_set_option("encode_path", ...): every path gives the same bits
  const int sms = ssdx_sm_count(anchors->device);
  if (opt != SSDX_ENC_TILE_FUSED && opt != SSDX_ENC_TILE_TILED &&
      ssdx_encode_grid_applies(anchors, gmax, num_classes, iou_threshold)) {
    // the anchor table is a union of regular grids (every DefaultBoundingBoxes output): closed-form windows
    const bool fits = ssdx_encode_grid_fused_fits(anchors, gmax, num_classes);
    // one (persistent) CTA per image whenever its tables fit: 15 us against 26 us for the split driver at batch 64,
    // 61 against 103 us at batch 1024; the split driver is for tables beyond 32768 anchors / 256 GT boxes
    const bool fused = fits && opt != SSDX_ENC_GRID_SPLIT;
    return ssdx_encode_grid_launch(anchors, fused, gt_xyxy, gt_label, gt_count, batch, gmax, num_classes,
                                   iou_threshold, std_host, out_labels, out_offsets, out_match, workspace, s);
  }
  if (use_fused(A, gmax, num_classes, batch, sms, opt)) {
    const size_t smem = fused_smem_bytes(A, gmax, num_classes);
    SSDX_CUDA(cudaFuncSetAttribute(ssdx_encode_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ssdx_encode_fused_kernel<<<batch, kFusedThreads, smem, s>>>(
        anchors->m_box, anchors->m_area, anchors->m_idx, anchors->t_bbox, anchors->t_area, anchors->n_tiles,
        anchors->enc_cent, (const float4*)gt_xyxy, gt_label, gt_count, A, gmax,
        num_classes, iou_threshold, std_host[0], std_host[1], std_host[2], std_host[3], out_labels,
        (float4*)out_offsets, out_match);
    SSDX_LAUNCH_CHECK("ssdx_encode_fused_kernel");
    return SSDX_OK;
  }

  const size_t best_bytes = ssdx_align_up((size_t)batch * gm * sizeof(u64), 256);
  u64* gt_best = (u64*)workspace;
  int* anchor_g2 = (int*)((char*)workspace + best_bytes);
  SSDX_CUDA(cudaMemsetAsync(gt_best, 0, (size_t)batch * gm * sizeof(u64), s));
  SSDX_CUDA(cudaMemsetAsync(anchor_g2, 0xFF, (size_t)batch * A * sizeof(int), s));  // -1: below the threshold
  dim3 g0((gm + kTiledThreads / 32 - 1) / (kTiledThreads / 32), batch);
  ssdx_encode_seed_kernel<<<g0, kTiledThreads, 0, s>>>(anchors->m_box, anchors->m_area, anchors->m_idx, anchors->t_bbox,
                                                       anchors->t_area, anchors->n_tiles, (const float4*)gt_xyxy,
                                                       gt_count, gm, gt_best);
  SSDX_LAUNCH_CHECK("ssdx_encode_seed_kernel");
  dim3 g1((anchors->n_tiles + kTiledTilesPerCta - 1) / kTiledTilesPerCta, batch);
  ssdx_encode_match_kernel<<<g1, kTiledThreads, 0, s>>>(anchors->m_box, anchors->m_area, anchors->m_idx,
                                                        anchors->t_bbox, anchors->t_area, anchors->n_tiles,
                                                        (const float4*)gt_xyxy, gt_count, A, gm, iou_threshold,
                                                        gt_best, anchor_g2);
  SSDX_LAUNCH_CHECK("ssdx_encode_match_kernel");
  dim3 g2((A + kEmitAnchors - 1) / kEmitAnchors, batch);
  ssdx_encode_emit_kernel<<<g2, kEmitThreads, 0, s>>>(anchors->enc_cent, (const float4*)gt_xyxy, gt_label, gt_count,
                                                      A, gm, num_classes, std_host[0], std_host[1], std_host[2],
                                                      std_host[3], gt_best, anchor_g2, out_labels,
                                                      (float4*)out_offsets, out_match);
  SSDX_LAUNCH_CHECK("ssdx_encode_emit_kernel");
  return SSDX_OK;
}
