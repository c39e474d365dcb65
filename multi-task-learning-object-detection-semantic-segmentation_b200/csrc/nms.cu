// Per-class confidence-filtered NMS + cross-class merge
// (reference: layers.py:127-168 -> tf.image.combined_non_max_suppression, q = 1, clip_boxes = False).
//
// Per-class pipeline (max_per_class < max_total; 4 launches, stream-ordered, all state in the caller's workspace;
// caps of at most 32 boxes per class take the one-warp-per-segment kernel of section 2w instead of steps 2-4):
//   1. select  : one CTA per (image, class group).  Scores are read coalesced (128-bit when C == 4).
//                A 1/16 sample of the anchors gives, per class, the key of its m-th largest sample
//                (warp-level bisection, REDUX only); a second pass appends every candidate with
//                score > thr and key >= that threshold to a per-(image, class) list: the top ~400 of
//                9600 candidates, unordered, as 64-bit composites (score key << 32 | ~anchor).
//   2. fast    : one CTA per (image, class).  Sorts the list (bitonic, shared memory), then greedy
//                suppression in blocks of 64 sorted candidates: each candidate is tested against the kept
//                boxes (4 threads per candidate), the 64x64 overlap matrix of the block is built, and one
//                warp resolves the block with a parallel fix-point (a candidate is kept once every earlier
//                overlapping candidate is known suppressed; suppressed once one is known kept) - a few
//                rounds instead of one serial step per box.  If the list overflowed, or ran dry before
//                min(max_per_class, A) boxes were kept while candidates remain, the segment is flagged.
//   3. generic : same grid, exits immediately unless flagged.  Exact for any input: keys of all
//                candidates in shared memory, repeated top-k' selection by 64-bit bisection on the composite
//                key (all keys distinct, so massive score ties terminate and are visited lower anchor first),
//                bitonic sort, blocked suppression with a cooperative warp sweep.
//   4. merge   : one CTA per image: rank-merges the C sorted pick lists (score desc; ties: lower class,
//                lower anchor), keeps max_total rows and writes [label, prob, xmin, ymin, xmax, ymax]
//                (layers.py:155-162), zero padding, valid counts and source anchor indices.
//
// Merged pipeline (2 launches, no fallback behind it), used when max_per_class >= max_total (the per-class cap can
// then never bind before the image has max_total rows), C <= 256 and A <= 2^24:
//   The K = max_total output rows are the first K boxes selected when the candidates of ALL classes are visited in
//   one descending order (score, then lower class, then lower anchor) and each one is tested only against the
//   boxes already selected for its own class: within a class that is exactly the per-class greedy order, and the
//   visiting order is the merge order, so everything after the K-th selection can never reach the output.  One
//   list, one sort and one greedy pass per image instead of one per (image, class), and no merge kernel.
//   1m. select_all : one CTA per image (256 threads; 512 / 1024 when the batch leaves SMs idle): threshold from a
//                    1/16 sample of all classes (bisection over the per-thread sample maxima), then one streaming
//                    pass with a single compare per score; ~400 composites
//                    (score key << 32 | ~class << 24 | ~anchor) per image.
//   2m. fast_all   : one CTA per image, launched as a programmatic dependent of select_all: a tier above 512
//                    candidates is trimmed by a bisection over the score keys; sort by counting (256 bins over the
//                    score keys + exact rank inside the bin; bitonic network above 512 keys); the sorted
//                    candidates are taken in chunks of up to 256: box gather, test against the boxes kept by
//                    earlier chunks, the chunk against itself as 32 x 32 tiles of a lower-triangular bit matrix
//                    (all warps, no barrier inside; fp16 prefilter boxes stored as column pairs, class folded in as
//                    a coordinate shift; class test and exact fp32 IoU on the pairs that pass), one warp resolves
//                    the chunk from the bit rows with a parallel fix-point, all warps copy the kept boxes; rows
//                    staged in shared memory.  A list that runs dry - or that overflowed, or that holds thousands of
//                    equal scores - is refilled in bands of the full 64-bit composite order (score key, class,
//                    anchor), narrowed down to a single composite if need be: the kernel never hands an image over
//                    to the generic kernel.  (ssdx_set_option("nms_path", "blocked"): the round-1 form of the greedy
//                    pass, 64 candidates at a time against the kept boxes.)
//
// IoU follows TF's IOU(): min/max-normalised corners, area = (y1-y0)*(x1-x0), 0 if either area <= 0,
// inter / ((area_i + area_j) - inter) with IEEE division, suppressed iff IoU > iou_thr (strict).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

typedef unsigned long long u64;

constexpr int kSelThreads = 256;
constexpr int kSelGroup = 16;      // classes per select CTA
constexpr int kSampleStride = 16;  // 1/16 of the anchors are sampled for the threshold estimate
constexpr int kTierRank = 20;      // first tier ends at the 20th largest sample => ~320 +- 70 listed candidates
constexpr int kInfo = 4;           // per segment: listed (tier 0), total candidates, key of the tier-0 threshold, pad
constexpr int kCap = 1024;         // list capacity per (image, class) tier
constexpr int kRefillTarget = 320; // candidates aimed at by each lazy refill of the fast kernel
constexpr int kMaxRefills = 24;
constexpr int kFastThreads = 256;
constexpr int kFastMaxSel = 512;   // fast path keeps at most this many boxes per class
constexpr int kBlk = 64;
constexpr int kGenThreads = 512;
constexpr int kChunk = 512;
constexpr int kMergeThreads = 256;
#ifndef SSDX_FAST_MINB
#define SSDX_FAST_MINB 7
#endif
#ifndef SSDX_ALL_RANK
#define SSDX_ALL_RANK 26
#endif
constexpr int kAllRank = SSDX_ALL_RANK;       // merged select: threshold at the 26th largest sample => ~416 +- 80 listed
constexpr int kAllRefillTarget = 384;
constexpr int kChunkTiles = 36;  // 32 x 32 tiles of the lower triangle of a 256-candidate chunk

#ifdef SSDX_PHASE_TIMERS
static __device__ unsigned long long g_phase_cycles[32];
#endif

__device__ __forceinline__ unsigned score_to_key(float s) {
  const unsigned u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_score(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
__device__ __forceinline__ u64 make_comp(unsigned key, int idx) {
  return ((u64)key << 32) | (u64)(0xFFFFFFFFu - (unsigned)idx);
}
__device__ __forceinline__ int comp_idx(u64 c) { return (int)(0xFFFFFFFFu - (unsigned)(c & 0xFFFFFFFFull)); }
// merged order: score descending, then lower class, then lower anchor (class < 256, anchor < 2^24)
__device__ __forceinline__ u64 make_comp_all(unsigned key, int cls, int idx) {
  return ((u64)key << 32) | ((u64)(0xFFu - (unsigned)cls) << 24) | (u64)(0xFFFFFFu - (unsigned)idx);
}
__device__ __forceinline__ int comp_all_cls(u64 c) { return (int)(0xFFu - (unsigned)((c >> 24) & 0xFFull)); }
__device__ __forceinline__ int comp_all_idx(u64 c) { return (int)(0xFFFFFFu - (unsigned)(c & 0xFFFFFFull)); }

// Optional per-class factor applied to every score as it is read (layers.SegmentationSuppression, layers.py:209:
// probabilities * is_class_segmented): one IEEE multiply, so the products equal the reference's; x * 1.0f == x.
__device__ __forceinline__ float class_factor(const float* __restrict__ cscale, int c) {
  return cscale ? __ldg(cscale + c) : 1.0f;
}
__device__ __forceinline__ float4 class_factor4(const float* __restrict__ cscale) {
  return cscale ? make_float4(__ldg(cscale), __ldg(cscale + 1), __ldg(cscale + 2), __ldg(cscale + 3))
                : make_float4(1.f, 1.f, 1.f, 1.f);
}
__device__ __forceinline__ float4 mul4(const float4 a, const float4 b) {
  return make_float4(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z), __fmul_rn(a.w, b.w));
}

// score of anchor a for class c of one image (cls_scores = image base + c): the stored probability, or - LOGITS, four
// aligned classes - the soft-max of the anchor's row of logits, recomputed as everywhere else (ssdx_softmax4)
template <bool LOGITS>
__device__ __forceinline__ float class_score(const float* __restrict__ cls_scores, const int a, const int C, const int c,
                                             const float cf) {
  if (LOGITS) {
    const float4 p = ssdx_softmax4(__ldg(reinterpret_cast<const float4*>(cls_scores - c) + a));
    return __fmul_rn(c == 0 ? p.x : (c == 1 ? p.y : (c == 2 ? p.z : p.w)), cf);
  }
  return __fmul_rn(__ldg(cls_scores + (size_t)a * C), cf);
}

// box normalised as TF does: (y0, x0, y1, x1) with y0<=y1, x0<=x1
__device__ __forceinline__ float4 normalise_box(const float4 b) {
  return make_float4(fminf(b.x, b.z), fminf(b.y, b.w), fmaxf(b.x, b.z), fmaxf(b.y, b.w));
}
__device__ __forceinline__ float box_area(const float4 n) {
  return __fmul_rn(__fsub_rn(n.z, n.x), __fsub_rn(n.w, n.y));
}
// fl(x - y) > 0  <=>  x > y, so "intersection height and width > 0" is four compares on the normalised corners
__device__ __forceinline__ bool boxes_overlap(const float4 a, const float4 b) {
  return (a.z > b.x) & (b.z > a.x) & (a.w > b.y) & (b.w > a.y);  // predicate AND chain, no branches
}
// Conservative half-precision copy of a normalised box for the overlap prefilter: minima rounded down, maxima rounded
// up, so "a.max > b.min" in fp32 implies the same in fp16 (never a false negative); 8 bytes and two packed compares
// per pair instead of 16 bytes and four.  Pairs that pass are decided by the exact fp32 test.
struct __align__(8) HBox {
  __half2 mn, mx;
};
__device__ __forceinline__ HBox make_hbox(const float4 n) {
  HBox h;
  h.mn = __halves2half2(__float2half_rd(n.x), __float2half_rd(n.y));
  h.mx = __halves2half2(__float2half_ru(n.z), __float2half_ru(n.w));
  return h;
}
__device__ __forceinline__ bool hboxes_may_overlap(const HBox a, const HBox b) {
  return __hbgt2(a.mx, b.mn) && __hbgt2(b.mx, a.mn);
}
// Prefilter copy with the class folded in: boxes of class c are shifted by unit * (c & 15, c >> 4) before the directed
// roundings (fp32 add and fp16 conversion both round minima down and maxima up; unit is a power of two, so the
// shifts are exact), so that boxes of the same class keep "exact overlap implies prefilter overlap" while boxes of
// different classes whose coordinates stay inside (-unit / 2, unit / 2) never pass it.  Pairs that pass are still
// decided by the class test and the exact fp32 IoU, so a unit that is too small (or a shift that saturates the
// fp16 range) only costs exact tests.
__device__ __forceinline__ HBox make_hbox_cls(const float4 n, const int cls, const float unit) {
  const float oy = __fmul_rn(unit, (float)(cls & 15)), ox = __fmul_rn(unit, (float)(cls >> 4));
  HBox h;
  h.mn = __halves2half2(__float2half_rd(__fadd_rd(n.x, oy)), __float2half_rd(__fadd_rd(n.y, ox)));
  h.mx = __halves2half2(__float2half_ru(__fadd_ru(n.z, oy)), __float2half_ru(__fadd_ru(n.w, ox)));
  return h;
}
// power of two above twice the largest coordinate magnitude `ext` (>= 0, finite or not), within [2^-8, 2^14]
__device__ __forceinline__ float class_unit(const float ext) {
  int e = (int)((__float_as_uint(ext) >> 23) & 0xFFu) - 127 + 2;  // ext < 2^(e - 1)
  e = e < -8 ? -8 : (e > 14 ? 14 : e);
  return __uint_as_float((unsigned)(e + 127) << 23);
}
// The prefilter boxes of a chunk are stored for the tile loop as column pairs: entry 16 * G + p holds candidates
// 32 * G + p (low halves) and 32 * G + 16 + p (high halves), one half2 per corner, so that one packed compare tests
// one corner of two columns against a row and the two half results ARE the two columns (no "both halves" reduction):
// 3.5 instructions per pair instead of 6.5.
struct __align__(16) HPair {
  __half2 ymn, ymx, xmn, xmx;
};
__device__ __forceinline__ void store_hpair(HPair* pairs, const int i, const HBox h) {
  __half* f = reinterpret_cast<__half*>(pairs + ((i >> 5) * 16 + (i & 15))) + ((i >> 4) & 1);
  f[0] = __low2half(h.mn);   // ymn
  f[2] = __low2half(h.mx);   // ymx
  f[4] = __high2half(h.mn);  // xmn
  f[6] = __high2half(h.mx);  // xmx
}
__device__ __forceinline__ HBox load_hbox(const HPair* pairs, const int i) {
  const __half* f = reinterpret_cast<const __half*>(pairs + ((i >> 5) * 16 + (i & 15))) + ((i >> 4) & 1);
  HBox h;
  h.mn = __halves2half2(f[0], f[4]);
  h.mx = __halves2half2(f[2], f[6]);
  return h;
}
// tile t of the lower triangle (t = g (g + 1) / 2 + jb, jb <= g) -> row group g
__constant__ unsigned char kTileRowGroup[36] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5,
                                                5, 5, 5, 6, 6, 6, 6, 6, 6, 6, 7, 7, 7, 7, 7, 7, 7, 7};

// exact TF IoU test for a pair that overlaps
__device__ __forceinline__ bool iou_exceeds_overlapping(const float4 a, const float aa, const float4 b, const float ab,
                                                        const float thr) {
  if (!(aa > 0.0f) || !(ab > 0.0f)) return false;  // TF: IoU = 0 when either area <= 0
  const float ih = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
  const float iw = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
  const float inter = __fmul_rn(fmaxf(ih, 0.0f), fmaxf(iw, 0.0f));
  const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
  return iou > thr;
}
// the same decision without branches (two of these in flight per thread in the tile loop): a pair that does not
// overlap has inter = 0 (or NaN from 0 * inf), whose quotient never exceeds thr >= 0
__device__ __forceinline__ bool iou_exceeds_flat(const float4 a, const float aa, const float4 b, const float ab,
                                                 const float thr) {
  const float ih = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
  const float iw = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
  const float inter = __fmul_rn(fmaxf(ih, 0.0f), fmaxf(iw, 0.0f));
  const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
  return (aa > 0.0f) & (ab > 0.0f) & (iou > thr);
}
__device__ __forceinline__ bool iou_exceeds(const float4 a, const float aa, const float4 b, const float ab,
                                            const float thr) {
  // disjoint boxes: IoU == 0 never exceeds thr >= 0
  return boxes_overlap(a, b) && iou_exceeds_overlapping(a, aa, b, ab, thr);
}

// in-place bitonic sort (descending) of N (power of two) u64 in shared memory by the whole CTA; N is a
// compile-time constant so the stage/partner arithmetic folds into immediates
template <int THREADS, int N>
__device__ __forceinline__ void bitonic_sort_fixed(u64* v, const int tid) {
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int p0 = 0; p0 < (N >> 1); p0 += THREADS) {
        const int p = p0 + tid;
        if ((N >> 1) >= THREADS || p < (N >> 1)) {
          const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
          const int q = i | j;
          const u64 x = v[i], y = v[q];
          const bool up = ((i & k) != 0);
          if ((x < y) != up && x != y) {
            v[i] = y;
            v[q] = x;
          }
        }
      }
      __syncthreads();
    }
  }
}
template <int THREADS>
__device__ __forceinline__ int bitonic_sort_desc(u64* v, const int n, const int tid) {
  // pads [n, n_sort) with zeros (sorted last) and sorts; returns n_sort.  Uniform control flow.
  int n_sort = 64;
  while (n_sort < n) n_sort <<= 1;
  for (int i = n + tid; i < n_sort; i += THREADS) v[i] = 0ull;
  __syncthreads();
  switch (n_sort) {
    case 64: bitonic_sort_fixed<THREADS, 64>(v, tid); break;
    case 128: bitonic_sort_fixed<THREADS, 128>(v, tid); break;
    case 256: bitonic_sort_fixed<THREADS, 256>(v, tid); break;
    case 512: bitonic_sort_fixed<THREADS, 512>(v, tid); break;
    default: bitonic_sort_fixed<THREADS, 1024>(v, tid); break;
  }
  return n_sort;
}

// NOTE: helpers that take shared-memory pointers deliberately carry no __restrict__: with it nvcc moved the
// caller's direct shared-memory accesses across the inlined exchange stages.
// Bitonic sort (descending) of 256*EPT u64 held EPT per thread by a 256-thread CTA: element e = r*256 + tid.
// Partner distance j < 32 -> warp shuffle, j >= 256 -> register exchange inside the thread, 32 <= j <= 128 ->
// shared-memory exchange (double-buffered: one barrier per such stage).  9 barriers instead of 45 for 512 keys.
template <int EPT, int TPB = 256>  // TPB threads hold EPT keys each: key e = r * TPB + tid
__device__ __forceinline__ void bitonic_sort_regs(u64 (&v)[EPT], u64* buf0, u64* buf1,
                                                  const int tid, const bool active = true) {
  constexpr int N = TPB * EPT;
  int flip = 0;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= TPB) {
        constexpr int dummy = 0;
        (void)dummy;
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
          const int rp = r ^ (j / TPB);
          if (rp > r) {  // the pair (r, rp) lives in this thread; r is the lower index
            const int e = r * TPB + tid;
            const bool desc = ((e & k) == 0);
            const u64 a = v[r], b = v[rp];
            const bool swap = desc ? (a < b) : (a > b);
            v[r] = swap ? b : a;
            v[rp] = swap ? a : b;
          }
        }
      } else if (j < 32) {
        if (active) {  // (whole warps are active or not)
#pragma unroll
          for (int r = 0; r < EPT; ++r) {
            const int e = r * TPB + tid;
            const u64 p = __shfl_xor_sync(0xFFFFFFFFu, v[r], j);
            const bool take_max = (((e & j) == 0) == ((e & k) == 0));
            v[r] = ((v[r] > p) == take_max) ? v[r] : p;  // keys are distinct or equal: one compare, one select
          }
        }
      } else {
        u64* buf = flip ? buf1 : buf0;
        flip ^= 1;
        if (active) {
#pragma unroll
          for (int r = 0; r < EPT; ++r) buf[r * TPB + tid] = v[r];
        }
        __syncthreads();
        if (active) {
#pragma unroll
          for (int r = 0; r < EPT; ++r) {
            const int e = r * TPB + tid;
            const u64 p = buf[e ^ j];
            const bool take_max = (((e & j) == 0) == ((e & k) == 0));
            v[r] = ((v[r] > p) == take_max) ? v[r] : p;
          }
        }
      }
    }
  }
}

// sort s_list[0..n) descending for the fast kernels; uniform control flow.  TPB = 256: threads 0..255 hold the keys.
// TPB = 512 (CTAs of at least 512 threads): lists above 256 keys are spread over 512 threads.  Threads that hold no
// keys only take part in the barriers.
template <int TPB = 256>
__device__ __forceinline__ void fast_sort(u64* s_list, u64* s_tmp, const int n, const int tid) {
  if (n <= 256) {
    const bool active = tid < 256;
    u64 v[1];
    v[0] = (active && tid < n) ? s_list[tid] : 0ull;
    __syncthreads();
    bitonic_sort_regs<1>(v, s_list, s_tmp, tid, active);
    __syncthreads();
    if (active) s_list[tid] = v[0];
  } else if (n <= 512) {
    constexpr int EPT = 512 / TPB;
    const bool active = tid < TPB;
    u64 v[EPT];
#pragma unroll
    for (int r = 0; r < EPT; ++r) v[r] = (active && r * TPB + tid < n) ? s_list[r * TPB + tid] : 0ull;
    __syncthreads();
    bitonic_sort_regs<EPT, TPB>(v, s_list, s_tmp, tid, active);
    __syncthreads();
    if (active) {
#pragma unroll
      for (int r = 0; r < EPT; ++r) s_list[r * TPB + tid] = v[r];
    }
  } else {
    constexpr int EPT = 1024 / TPB;
    const bool active = tid < TPB;
    u64 v[EPT];
#pragma unroll
    for (int r = 0; r < EPT; ++r) v[r] = (active && r * TPB + tid < n) ? s_list[r * TPB + tid] : 0ull;
    __syncthreads();
    bitonic_sort_regs<EPT, TPB>(v, s_list, s_tmp, tid, active);
    __syncthreads();
    if (active) {
#pragma unroll
      for (int r = 0; r < EPT; ++r) s_list[r * TPB + tid] = v[r];
    }
  }
  __syncthreads();
}

// Sort s_list[0..n), n <= 512, descending, by counting instead of a network (fast_all, ALGO 1): the keys are binned
// by score key (256 bins over [min, max], a monotone map), a scan over the bins from the top gives every bin its base,
// the keys are scattered bin by bin into s_tmp, and every key then counts the larger keys of its own bin: base + count
// is its exact rank (the composites are distinct).  ~60 instructions per key and 6 barriers where the 512-key bitonic
// network spends ~380 per key and 11; thousands of equal score keys (one bin) only cost the quadratic count.
// s_bins: 512 words of scratch (counts, then bases), s_mm: 2 words (min / max key).
template <int NT>
__device__ __forceinline__ void rank_sort_512(u64* s_list, u64* s_tmp, unsigned* s_bins, unsigned* s_mm, const int n,
                                              const int tid) {
  constexpr int EPT = (512 + NT - 1) / NT;
  const int lane = tid & 31;
  u64 v[EPT];
  unsigned kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll
  for (int r = 0; r < EPT; ++r) {
    const int e = r * NT + tid;
    v[r] = (e < n) ? s_list[e] : 0ull;
    if (e < n) {
      kmin = min(kmin, (unsigned)(v[r] >> 32));
      kmax = max(kmax, (unsigned)(v[r] >> 32));
    }
  }
  if (tid < 256) s_bins[tid] = 0u;
  if (tid == 0) {
    s_mm[0] = 0xFFFFFFFFu;
    s_mm[1] = 0u;
  }
  __syncthreads();
  kmin = __reduce_min_sync(0xFFFFFFFFu, kmin);
  kmax = __reduce_max_sync(0xFFFFFFFFu, kmax);
  if (lane == 0) {
    atomicMin(&s_mm[0], kmin);
    atomicMax(&s_mm[1], kmax);
  }
  __syncthreads();
  kmin = s_mm[0];
  kmax = s_mm[1];
  // bin = floor((key - min) * 256 / (max - min + 1)) in fp32: conversion, product and truncation are all monotone
  const float inv = __fdividef(256.0f, (float)(kmax - kmin) + 1.0f);
  int bin[EPT], pos[EPT];
#pragma unroll
  for (int r = 0; r < EPT; ++r) {
    bin[r] = 0;
    pos[r] = 0;
    if (r * NT + tid < n) {
      bin[r] = min(255, (int)((float)((unsigned)(v[r] >> 32) - kmin) * inv));
      pos[r] = (int)atomicAdd(&s_bins[bin[r]], 1u);
    }
  }
  __syncthreads();
  if (tid < 32) {  // bases: exclusive scan of the counts from the highest bin down; lane l owns bins 255 - 8 l - u
    unsigned c[8], sum = 0u;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      c[u] = s_bins[255 - 8 * lane - u];
      sum += c[u];
    }
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += y;
    }
    unsigned base = incl - sum;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s_bins[256 + 255 - 8 * lane - u] = base;
      base += c[u];
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < EPT; ++r)
    if (r * NT + tid < n) s_tmp[s_bins[256 + bin[r]] + pos[r]] = v[r];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < EPT; ++r) {
    if (r * NT + tid < n) {
      const int base = (int)s_bins[256 + bin[r]], cnt = (int)s_bins[bin[r]];
      int larger = 0;
      for (int i = 0; i < cnt; ++i) larger += s_tmp[base + i] > v[r];
      s_list[base + larger] = v[r];
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------ 1. select
// info per (image, class): [0] listed candidates (may exceed kCap), [2] threshold key (every candidate with key >= it
// is listed).  Same scheme as the merged select below, per class: threshold = the kTierRank-th largest of the
// per-thread maxima of a 1/16 sample, then one streaming pass with one compare per score and marked hits.
template <bool VEC4, bool LOGITS>  // LOGITS (VEC4 only): `scores` holds raw logits, see ssdx_nms_select_all_kernel
__global__ void __launch_bounds__(kSelThreads)
    ssdx_nms_select_kernel(const float* __restrict__ scores, const float* __restrict__ cscale, const int A,
                           const int C, const float score_thr, const int tier_rank, u64* __restrict__ lists,
                           int* __restrict__ info, int* __restrict__ flags) {
  __shared__ unsigned s_max[kSelGroup][kSelThreads];  // per class: per-thread maximum of the sampled keys
  __shared__ unsigned s_T[kSelGroup];
  __shared__ float s_tlo[kSelGroup];
  __shared__ int s_cnt[kSelGroup];
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * kSelGroup;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) flags[0] = 0;  // worklist of the generic kernel
  const int cg = min(kSelGroup, C - c0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NS = (A + kSampleStride - 1) / kSampleStride;
  const float* img = scores + (size_t)b * A * C;
  if (tid < kSelGroup) s_cnt[tid] = 0;
  for (int c = 0; c < cg; ++c) s_max[c][tid] = 0u;  // own column only: no barrier needed before the sample pass
  // ---- sample ----
  for (int k = tid; k < NS; k += kSelThreads) {
    int a = k * kSampleStride + ((k * 7) & (kSampleStride - 1));
    a = a < A ? a : A - 1;
    if (VEC4) {
      float4 raw = __ldg(reinterpret_cast<const float4*>(img) + a);
      if (LOGITS) raw = ssdx_softmax4(raw);
      const float4 s = mul4(raw, class_factor4(cscale));
      const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (sv[c] > score_thr) s_max[c][tid] = max(s_max[c][tid], score_to_key(sv[c]));
    } else {
      for (int c = 0; c < cg; ++c) {
        const float s = __fmul_rn(__ldg(img + (size_t)a * C + c0 + c), class_factor(cscale, c0 + c));
        if (s > score_thr) s_max[c][tid] = max(s_max[c][tid], score_to_key(s));
      }
    }
  }
  __syncthreads();
  // ---- threshold: one warp per class bisects the 256 maxima held in registers ----
  for (int c = warp; c < cg; c += kSelThreads / 32) {
    unsigned v[kSelThreads / 32];
    int nz = 0;
    unsigned kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll
    for (int r = 0; r < kSelThreads / 32; ++r) {
      v[r] = s_max[c][r * 32 + lane];
      nz += v[r] != 0u;
      kmax = max(kmax, v[r]);
      kmin = min(kmin, v[r] != 0u ? v[r] : 0xFFFFFFFFu);
    }
    nz = __reduce_add_sync(0xFFFFFFFFu, nz);
    kmax = __reduce_max_sync(0xFFFFFFFFu, kmax);
    kmin = __reduce_min_sync(0xFFFFFFFFu, kmin);
    unsigned T = 1u;  // list every candidate
    // a class is thresholded only when it is dense (>= kTierRank threads saw a candidate): a sparse class is listed
    // whole, so that it never needs a refill scan; the rank of the threshold follows the per-class cap
    if (A > kCap && nz >= kTierRank) {
      unsigned lo = kmin;       // count(max >= kmin) = nz >= rank
      unsigned hi = kmax + 1u;  // count = 0
      while (hi - lo > 1u) {
        const unsigned mid = lo + ((hi - lo) >> 1);
        int cnt = 0;
#pragma unroll
        for (int r = 0; r < kSelThreads / 32; ++r) cnt += v[r] >= mid;
        cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
        if (cnt >= tier_rank) {
          lo = mid;
          if (cnt <= tier_rank + 3) break;  // any threshold in this band will do
        } else {
          hi = mid;
        }
      }
      T = lo;
      if (lo == kmin && nz <= tier_rank + 3) T = 1u;  // every thread maximum is in the band: list every candidate
    }
    if (T == 0x80000000u) T = 0x7FFFFFFFu;  // +0.0 -> -0.0: keeps "s >= t_lo" and "key >= T" the same set
    if (lane == 0) {
      s_T[c] = T;
      // listed  <=>  s > score_thr and key(s) >= T  <=>  s >= t_lo   (one compare per score)
      const float tf = (T <= 1u) ? -INFINITY : key_to_score(T);
      s_tlo[c] = (tf > score_thr) ? tf : nextafterf(score_thr, INFINITY);
    }
  }
  __syncthreads();
  // ---- one streaming pass; hits are only marked, then appended in a compact loop ----
  u64* img_lists = lists + ((size_t)b * C + c0) * kCap;
  if (VEC4) {
    const float t0 = s_tlo[0], t1 = s_tlo[1], t2 = s_tlo[2], t3 = s_tlo[3];
    const float4 cf4 = class_factor4(cscale);
    constexpr int kBatch = 8;  // independent 128-bit streaming loads in flight per thread
    for (int a0 = tid; a0 < A; a0 += kSelThreads * kBatch) {
      float4 sb[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int a = a0 + u * kSelThreads;
        sb[u] = (a < A) ? __ldcs(reinterpret_cast<const float4*>(img) + a)
                        : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      }
      if (LOGITS) {
#pragma unroll
        for (int u = 0; u < kBatch; ++u)
          if (a0 + u * kSelThreads < A) sb[u] = ssdx_softmax4(sb[u]);
      }
      unsigned hits = 0u;  // bit 4*u + c
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        sb[u] = mul4(sb[u], cf4);
        hits |= (sb[u].x >= t0 ? 1u : 0u) << (4 * u);
        hits |= (sb[u].y >= t1 ? 2u : 0u) << (4 * u);
        hits |= (sb[u].z >= t2 ? 4u : 0u) << (4 * u);
        hits |= (sb[u].w >= t3 ? 8u : 0u) << (4 * u);
      }
      while (hits != 0u) {  // the score is re-read (L1/L2 hit) instead of indexing the register batch dynamically
        const int bit = __ffs(hits) - 1;
        hits &= hits - 1u;
        const int a = a0 + (bit >> 2) * kSelThreads, c = bit & 3;
        float s;
        if (LOGITS) {
          const float4 p = ssdx_softmax4(__ldg(reinterpret_cast<const float4*>(img) + a));
          s = __fmul_rn(c == 0 ? p.x : (c == 1 ? p.y : (c == 2 ? p.z : p.w)), class_factor(cscale, c));
        } else {
          s = __fmul_rn(__ldg(img + (size_t)a * 4 + c), class_factor(cscale, c));
        }
        const int pos = atomicAdd(&s_cnt[c], 1);
        if (pos < kCap) img_lists[(size_t)c * kCap + pos] = make_comp(score_to_key(s), a);
      }
    }
  } else {
    for (int c = 0; c < cg; ++c) {
      const float t = s_tlo[c];
      const float cf = class_factor(cscale, c0 + c);
      for (int a = tid; a < A; a += kSelThreads) {
        const float s = __fmul_rn(__ldg(img + (size_t)a * C + c0 + c), cf);
        if (s >= t) {
          const int pos = atomicAdd(&s_cnt[c], 1);
          if (pos < kCap) img_lists[(size_t)c * kCap + pos] = make_comp(score_to_key(s), a);
        }
      }
    }
  }
  __syncthreads();
  if (tid < cg) {
    int* dst = info + ((size_t)b * C + c0 + tid) * kInfo;
    dst[0] = s_cnt[tid];
    dst[1] = 0;
    dst[2] = (int)s_T[tid];
    dst[3] = 0;
  }
}

// ------------------------------------------------------------------------------------------ 2. fast path
template <bool LOGITS>  // LOGITS: `scores` holds raw logits of 4 classes (the refill scans recompute the soft-max rows)
__global__ void __launch_bounds__(kFastThreads, 6)
    ssdx_nms_fast_kernel(const ssdx_box_src boxes, const float* __restrict__ scores,
                         const float* __restrict__ cscale, const u64* __restrict__ lists, const int* __restrict__ info, const int A, const int C,
                         const int limit, const float iou_thr, const float score_thr, u64* __restrict__ picks,
                         int* __restrict__ pick_cnt, int* __restrict__ flags) {
  __shared__ u64 s_list[kCap];
  __shared__ u64 s_tmp[kCap];  // second exchange buffer of the sort
  __shared__ float4 s_cbox[kFastThreads];
  __shared__ float s_carea[kFastThreads];
  __shared__ float4 s_sbox[kFastMaxSel];
  __shared__ float s_sarea[kFastMaxSel];
  __shared__ u64 s_row[kBlk];  // per candidate of the block: earlier candidates (j < i) that overlap it
  __shared__ int s_sup[kBlk];
  __shared__ int s_nsel, s_n;

  const int c = blockIdx.x, b = blockIdx.y;
  const size_t seg = (size_t)b * C + c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int* seg_info = info + seg * kInfo;
  int n = seg_info[0];
  if (n > kCap || limit > kFastMaxSel) {  // uniform
    if (tid == 0) flags[1 + atomicAdd(&flags[0], 1)] = (int)seg;  // hand the segment to the generic kernel
    return;
  }
  const size_t box_row0 = (size_t)b * A;
  const float* img_scores = scores + (size_t)b * A * C + c;
  const float cf = class_factor(cscale, c);
  u64* my_picks = picks + seg * limit;
  const u64* lst = lists + seg * kCap;
  PHASE_T0();
  for (int i = tid; i < n; i += kFastThreads) s_list[i] = lst[i];
  if (tid == 0) s_nsel = 0;
  __syncthreads();
  PHASE_ADD(0);  // list load

  int n_sel = 0, consumed = 0, scans = 0;
  bool failed = false;
  unsigned key_hi = (unsigned)seg_info[2];  // every candidate with key >= key_hi has been listed
  unsigned span = 0u;
  while (true) {
    // ---- sort the current tier (descending score, ties -> lower anchor) ----
    fast_sort(s_list, s_tmp, n, tid);
    PHASE_ADD(1);  // sort
    consumed += n;
    if (span == 0u) {  // key width that held the lower half of this tier -> width expected to hold kRefillTarget
      const unsigned k_mid = (n >= 2) ? (unsigned)(s_list[n >> 1] >> 32) : 0u;
      const unsigned k_bot = (n >= 2) ? (unsigned)(s_list[n - 1] >> 32) : 0u;
      const u64 w = (n >= 2) ? (u64)(k_mid - k_bot) * (2ull * kRefillTarget) / (u64)n : 0ull;
      span = (w == 0ull) ? 1024u : (w > 0x7FFFFFFFull ? 0x7FFFFFFFu : (unsigned)w);
    }

    for (int g0 = 0; g0 < n && n_sel < limit; g0 += kFastThreads) {
      const int gn = min(kFastThreads, n - g0);
      if (tid < gn) {  // gather the boxes of the next 256 sorted candidates
        const float4 nb = normalise_box(ssdx_load_box(boxes, box_row0, comp_idx(s_list[g0 + tid])));
        s_cbox[tid] = nb;
        s_carea[tid] = box_area(nb);
      }
      __syncthreads();
      PHASE_ADD(2);  // box gather
      for (int base = 0; base < gn && n_sel < limit; base += kBlk) {
        const int m = min(kBlk, gn - base);
        const int ci = tid >> 2, part = tid & 3;
        {
          bool hit = false;
          float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
          float ai = 0.f;
          if (ci < m) {
            bi = s_cbox[base + ci];
            ai = s_carea[base + ci];
          }
          if (ai > 0.0f) {  // a box without area is never suppressed
            // kept boxes, newest first (similar scores overlap most); 4 independent loads + 16 compares per trip,
            // the exact IoU only for the rare overlapping pair
            int s = n_sel - 1 - part;
            for (; s >= 12 && !hit; s -= 16) {
              const float4 b0 = s_sbox[s], b1 = s_sbox[s - 4], b2 = s_sbox[s - 8], b3 = s_sbox[s - 12];
              const bool o0 = boxes_overlap(bi, b0), o1 = boxes_overlap(bi, b1), o2 = boxes_overlap(bi, b2),
                         o3 = boxes_overlap(bi, b3);
              if (o0 | o1 | o2 | o3) {
                hit = (o0 && iou_exceeds_overlapping(bi, ai, b0, s_sarea[s], iou_thr)) ||
                      (o1 && iou_exceeds_overlapping(bi, ai, b1, s_sarea[s - 4], iou_thr)) ||
                      (o2 && iou_exceeds_overlapping(bi, ai, b2, s_sarea[s - 8], iou_thr)) ||
                      (o3 && iou_exceeds_overlapping(bi, ai, b3, s_sarea[s - 12], iou_thr));
              }
            }
            for (; s >= 0 && !hit; s -= 4) {
              const float4 b0 = s_sbox[s];
              if (boxes_overlap(bi, b0)) hit = iou_exceeds_overlapping(bi, ai, b0, s_sarea[s], iou_thr);
            }
          }
          const unsigned bal = __ballot_sync(0xFFFFFFFFu, hit);
          if (part == 0) s_sup[ci] = (ci >= m) || (((bal >> (lane & 28)) & 0xFu) != 0u);
          // earlier candidates j < ci that overlap ci; this thread tests j = 4*jj + part (interleaved slices keep
          // the four parts of a candidate on consecutive shared-memory addresses)
          unsigned lo = 0u, hi = 0u;
          if (ai > 0.0f && part < ci) {
            unsigned ov = 0u;
            const int nj = (ci - part + 3) >> 2;  // number of jj with 4*jj + part < ci
            const int nj4 = (nj + 3) & ~3;        // stale slots beyond m / ci are masked below
            for (int j0 = 0; j0 < nj4; j0 += 4) {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj)
                ov |= (unsigned)boxes_overlap(bi, s_cbox[base + 4 * (j0 + jj) + part]) << (j0 + jj);
            }
            if (nj < 16) ov &= (1u << nj) - 1u;
            while (ov) {  // exact test for the few overlapping pairs
              const int jj = __ffs(ov) - 1;
              ov &= ov - 1u;
              const int j = 4 * jj + part;
              if (iou_exceeds_overlapping(bi, ai, s_cbox[base + j], s_carea[base + j], iou_thr)) {
                if (j < 32) lo |= 1u << j; else hi |= 1u << (j - 32);
              }
            }
          }
          lo |= __shfl_xor_sync(0xFFFFFFFFu, lo, 1);
          hi |= __shfl_xor_sync(0xFFFFFFFFu, hi, 1);
          lo |= __shfl_xor_sync(0xFFFFFFFFu, lo, 2);
          hi |= __shfl_xor_sync(0xFFFFFFFFu, hi, 2);
          if (part == 0) s_row[ci] = ((u64)hi << 32) | (u64)lo;
        }
        __syncthreads();
        PHASE_ADD(3);  // pair tests
        if (warp == kFastThreads / 32 - 1) {  // the arbiter favours the highest warp id: keep the serial part there
          // lane l owns candidates l (A) and l+32 (B) of the block; all masks are pairs of 32-bit words
          const unsigned lt = (1u << lane) - 1u, me = 1u << lane;
          const u64 ra = s_row[lane], rb = s_row[lane + 32];
          const unsigned a_lo = (unsigned)ra & lt;  // earlier overlappers of A (all in the low word)
          const unsigned b_lo = (unsigned)rb, b_hi = (unsigned)(rb >> 32) & lt;  // earlier overlappers of B
          unsigned supp_lo = __ballot_sync(0xFFFFFFFFu, s_sup[lane] != 0);
          unsigned supp_hi = __ballot_sync(0xFFFFFFFFu, s_sup[lane + 32] != 0);
          unsigned kept_lo = 0u, kept_hi = 0u;
          // fix-point: the lowest undecided candidate is decided in every round
          while ((kept_lo | supp_lo) != 0xFFFFFFFFu || (kept_hi | supp_hi) != 0xFFFFFFFFu) {
            const bool und_a = !((kept_lo | supp_lo) & me), und_b = !((kept_hi | supp_hi) & me);
            const bool a_hit = (a_lo & kept_lo) != 0u;
            const bool b_hit = ((b_lo & kept_lo) | (b_hi & kept_hi)) != 0u;
            const bool a_free = (a_lo & ~supp_lo) == 0u;                          // every earlier overlapper suppressed
            const bool b_free = ((b_lo & ~supp_lo) | (b_hi & ~supp_hi)) == 0u;
            const unsigned nk_lo = __ballot_sync(0xFFFFFFFFu, und_a && !a_hit && a_free);
            const unsigned ns_lo = __ballot_sync(0xFFFFFFFFu, und_a && a_hit);
            const unsigned nk_hi = __ballot_sync(0xFFFFFFFFu, und_b && !b_hit && b_free);
            const unsigned ns_hi = __ballot_sync(0xFFFFFFFFu, und_b && b_hit);
            kept_lo |= nk_lo;
            supp_lo |= ns_lo;
            kept_hi |= nk_hi;
            supp_hi |= ns_hi;
          }
          const int room = limit - n_sel;
          const int n_lo = __popc(kept_lo);
          if (kept_lo & me) {
            const int r = __popc(kept_lo & lt);
            if (r < room) {
              s_sbox[n_sel + r] = s_cbox[base + lane];
              s_sarea[n_sel + r] = s_carea[base + lane];
              my_picks[n_sel + r] = s_list[g0 + base + lane];
            }
          }
          if (kept_hi & me) {
            const int r = n_lo + __popc(kept_hi & lt);
            if (r < room) {
              s_sbox[n_sel + r] = s_cbox[base + lane + 32];
              s_sarea[n_sel + r] = s_carea[base + lane + 32];
              my_picks[n_sel + r] = s_list[g0 + base + lane + 32];
            }
          }
          if (lane == 0) s_nsel = n_sel + min(room, n_lo + __popc(kept_hi));
        }
        __syncthreads();
        PHASE_ADD(4);  // block resolve
        n_sel = s_nsel;
      }
    }
    if (n_sel >= limit || key_hi <= 1u) break;  // done, or every candidate has been visited
    // ---- the tier ran dry and candidates remain: fetch the next key band [key_lo, key_hi) from the scores.
    //      The band width adapts to the density seen so far; an over-full band is narrowed and rescanned. ----
    bool got = false;
    while (!got && !failed) {
      if (key_hi <= 1u || ++scans > kMaxRefills) {
        failed = true;
        break;
      }
      const unsigned key_lo = (key_hi - 1u > span) ? key_hi - span : 1u;
      __syncthreads();
      if (tid == 0) s_n = 0;
      __syncthreads();
      for (int a0 = tid; a0 < A; a0 += kFastThreads * 8) {  // 8 independent loads in flight per thread
        float sv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int a = a0 + u * kFastThreads;
          sv[u] = (a < A) ? class_score<LOGITS>(img_scores, a, C, c, cf) : -INFINITY;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (sv[u] > score_thr) {  // false for the -inf padding
            const unsigned key = score_to_key(sv[u]);
            if (key >= key_lo && key < key_hi) {
              const int pos = atomicAdd(&s_n, 1);
              if (pos < kCap) s_list[pos] = make_comp(key, a0 + u * kFastThreads);
            }
          }
        }
      }
      __syncthreads();
      const int n_new = s_n;
      if (n_new > kCap) {
        if (span <= 1u) failed = true;  // more than kCap equal scores: cannot be split by a threshold
        span = max(span >> 2, 1u);
      } else {
        n = n_new;
        key_hi = key_lo;
        if (n_new < kRefillTarget / 8) span = (span > 0x1FFFFFFFu) ? 0x7FFFFFFFu : span << 2;
        else if (n_new < kRefillTarget / 2) span = (span > 0x3FFFFFFFu) ? 0x7FFFFFFFu : span << 1;
        else if (n_new > kRefillTarget * 2) span = max(span >> 1, 1u);
        got = n_new > 0 || key_lo <= 1u;
      }
    }
    if (failed) break;
    __syncthreads();
    PHASE_ADD(5);  // refill scans
  }
  PHASE_ADD(6);
  if (tid == 0) {
#ifdef SSDX_PHASE_TIMERS
    atomicAdd(&g_phase_cycles[8], 1ull);
    atomicAdd(&g_phase_cycles[9], (unsigned long long)scans);
    atomicAdd(&g_phase_cycles[10], (unsigned long long)consumed);
#endif
    const bool complete = !failed;
    pick_cnt[seg] = n_sel;
    if (!complete) flags[1 + atomicAdd(&flags[0], 1)] = (int)seg;
  }
}


// number of entries of a descending list whose score key is > k (strict) or >= k
__device__ __forceinline__ int count_before(const u64* lst, int n, unsigned k, bool inclusive) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const unsigned km = (unsigned)(lst[mid] >> 32);
    const bool before = inclusive ? (km >= k) : (km > k);
    if (before) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Rows of one image from its per-class pick lists (descending; list c at img_picks + c * stride, cnt[c] entries): the
// rank of an entry = entries before it in its own list + entries of the other classes with a higher score (equal
// scores: the lower class first).  Executed by `nthreads` threads (a CTA, or one warp of the small-caps kernel).
__device__ __forceinline__ void merge_image_rows(const ssdx_box_src boxes, const u64* img_picks, const int* cnt,
                                                 const int stride, const int b, const int A, const int C,
                                                 const int limit, const int K, float* __restrict__ out,
                                                 int* __restrict__ out_valid, int* __restrict__ out_idx, const int tid,
                                                 const int nthreads) {
  int total = 0;
  for (int c = 0; c < C; ++c) total += cnt[c];
  const int valid = min(total, K);
  float* o = out + (size_t)b * K * 6;
  for (int e = tid; e < C * limit; e += nthreads) {
    const int c = e / limit, j = e - c * limit;
    if (j >= cnt[c]) continue;
    const u64 cp = img_picks[(size_t)c * stride + j];
    const unsigned k = (unsigned)(cp >> 32);
    int rank = j;
    for (int c2 = 0; c2 < C && rank < K; ++c2)
      if (c2 != c) rank += count_before(img_picks + (size_t)c2 * stride, cnt[c2], k, c2 < c);
    if (rank < K) {
      const int idx = comp_idx(cp);
      const float4 bx = ssdx_load_box(boxes, (size_t)b * A, idx);  // raw (ymin, xmin, ymax, xmax)
      float* r = o + (size_t)rank * 6;
      r[0] = (float)c;
      r[1] = key_to_score(k);
      r[2] = bx.y;  // gather [1,0,3,2], layers.py:155
      r[3] = bx.x;
      r[4] = bx.w;
      r[5] = bx.z;
      if (out_idx) out_idx[(size_t)b * K + rank] = idx;
    }
  }
  for (int e = valid * 6 + tid; e < K * 6; e += nthreads) o[e] = 0.0f;
  if (out_idx)
    for (int e = valid + tid; e < K; e += nthreads) out_idx[(size_t)b * K + e] = -1;
  if (out_valid && tid == 0) out_valid[b] = valid;
}

// ------------------------------------------------------------------------------------------ 2w. per-class, small caps
// One WARP per (image, class) segment for per-class caps of at most 32 boxes (the reference's own inference setting is
// 4 per class / 10 per image, 03-...ipynb): a CTA per segment spends its life in barriers and dependent loads for a
// handful of boxes.  The warp holds up to SLOTS * 32 candidates in registers (composite, box, area) and runs the
// classic form of the same greedy rule: take the best candidate still alive, keep it, kill every candidate whose IoU
// with it exceeds the threshold - one round per KEPT box instead of one test per visited candidate.  When the list
// runs dry it fetches the next band of the composite order (score key, anchor) from the scores, tests the newcomers
// against the boxes kept so far and goes on; an over-full band is narrowed, down to a single candidate, so the kernel
// never hands a segment over (thousands of equal scores included).
constexpr int kWarpSegs = kFastThreads / 32;  // segments per CTA
constexpr int kWarpRefillTarget = 96;

// LOGITS: `scores` holds raw logits of 4 classes (band refills recompute the soft-max rows).  SLOTS candidates per lane
// in registers, MINB resident CTAs per SM: <8, 1> for caps up to 8 (few rounds: the whole listed tier in registers
// matters more than occupancy), <4, 4> above (many rounds: 32 resident warps per SM hide their latency; 20 / 50 / 0.3
// / 0.3 at batch 1024: 79 us against 104 us).
template <bool LOGITS, int SLOTS, int MINB>
__global__ void __launch_bounds__(kFastThreads, MINB)
    ssdx_nms_fast_warp_kernel(const ssdx_box_src boxes, const float* __restrict__ scores,
                              const float* __restrict__ cscale, const u64* __restrict__ lists,
                              const int* __restrict__ info, const int A, const int C, const int num_segments,
                              const int limit, const float iou_thr, const float score_thr, u64* __restrict__ picks,
                              int* __restrict__ pick_cnt, const int fuse_merge, const int K, float* __restrict__ out,
                              int* __restrict__ out_valid, int* __restrict__ out_idx) {
  // fuse_merge (the class count divides the segments of a CTA: every image lies inside one CTA): the picks stay in
  // shared memory and one warp per image writes the image's rows at the end - no pick lists in memory, no merge launch
  constexpr int kWarpSlots = SLOTS, kWarpCap = SLOTS * 32;
  __shared__ u64 s_stage[kWarpSegs][kWarpCap];  // refill staging, one row per warp
  __shared__ u64 s_picks[kWarpSegs][32];
  __shared__ int s_cnt[kWarpSegs];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const size_t seg = (size_t)blockIdx.x * kWarpSegs + warp;
  int n_sel = 0;
  if (seg < (size_t)num_segments) {  // warp-uniform
  const int b = (int)(seg / C), c = (int)(seg - (size_t)b * C);
  const size_t box_row0 = (size_t)b * A;
  const float* cls_scores = scores + (size_t)b * A * C + c;
  const float cf = class_factor(cscale, c);
  u64* my_picks = picks + seg * limit;
  u64* stage = s_stage[warp];

  u64 comp[kWarpSlots];
  float4 box[kWarpSlots];
  float area[kWarpSlots];
  unsigned alive = 0u;  // bit j: slot j holds a candidate that is neither kept nor suppressed
  float4 kbox = make_float4(0.f, 0.f, 0.f, 0.f);  // lane k holds the k-th kept box
  float karea = 0.0f;

  auto load_boxes = [&](const int n) {  // slots of the current tier -> boxes; entry i sits in lane i % 32, slot i / 32
#pragma unroll
    for (int j = 0; j < kWarpSlots; ++j) {
      const bool have = j * 32 + lane < n;
      box[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      area[j] = 0.0f;
      if (have) {
        box[j] = normalise_box(ssdx_load_box(boxes, box_row0, comp_idx(comp[j])));
        area[j] = box_area(box[j]);
        alive |= 1u << j;
      }
    }
  };
  // kill the candidates that the kept box (kb, ka) suppresses
  auto suppress_by = [&](const float4 kb, const float ka) {
#pragma unroll
    for (int j = 0; j < kWarpSlots; ++j)
      if (((alive >> j) & 1u) && iou_exceeds(box[j], area[j], kb, ka, iou_thr)) alive &= ~(1u << j);
  };

  const int n_listed = info[seg * kInfo + 0];
  const u64* lst = lists + seg * kCap;
  // composites >= list_floor are in the list (when it was not truncated); below it only the scores know
  const u64 list_floor = (u64)(unsigned)info[seg * kInfo + 2] << 32;
  const bool list_ok = n_listed <= kCap;
  u64 comp_hi;  // every candidate with composite >= comp_hi has been fetched
  u64 span;
#pragma unroll
  for (int j = 0; j < kWarpSlots; ++j) comp[j] = 0ull;
  if (n_listed <= kWarpCap) {  // the usual case: the whole listed tier fits the registers
    u64 cmin = 0xFFFFFFFFFFFFFFFFull, cmax = 0ull;
#pragma unroll
    for (int j = 0; j < kWarpSlots; ++j)
      if (j * 32 + lane < n_listed) {
        comp[j] = __ldg(lst + j * 32 + lane);
        cmin = comp[j] < cmin ? comp[j] : cmin;
        cmax = comp[j] > cmax ? comp[j] : cmax;
      }
    load_boxes(n_listed);
    // first refill band: the width that held kWarpRefillTarget candidates of this tier on average
    const unsigned hmin = __reduce_min_sync(0xFFFFFFFFu, (unsigned)(cmin >> 32));
    const unsigned hmax = __reduce_max_sync(0xFFFFFFFFu, (unsigned)(cmax >> 32));
    const u64 w = (n_listed >= 2 && hmax > hmin) ? ((u64)(hmax - hmin) << 32) / (u64)n_listed * kWarpRefillTarget
                                                 : (1024ull << 32);
    span = w < (1ull << 32) ? (1ull << 32) : w;
    comp_hi = list_floor;
  } else if (list_ok) {  // a longer list (a sparse class is listed whole): served in bands, from the list itself
    unsigned hmin = 0xFFFFFFFFu, hmax = 0u;
    for (int i = lane; i < n_listed; i += 32) {
      const unsigned k = (unsigned)(__ldg(lst + i) >> 32);
      hmin = min(hmin, k);
      hmax = max(hmax, k);
    }
    hmin = __reduce_min_sync(0xFFFFFFFFu, hmin);
    hmax = __reduce_max_sync(0xFFFFFFFFu, hmax);
    const u64 w = (hmax > hmin) ? ((u64)(hmax - hmin) << 32) / (u64)n_listed * kWarpRefillTarget : (1024ull << 32);
    span = w < (1ull << 32) ? (1ull << 32) : w;
    comp_hi = ((u64)hmax << 32) | 0xFFFFFFFFull;
    comp_hi = comp_hi == 0xFFFFFFFFFFFFFFFFull ? comp_hi : comp_hi + 1ull;  // just above the best listed candidate
  } else {  // truncated list (more than kCap candidates at or above the threshold key): bands from the scores
    span = (0xFFFFFFFFFFFFFFFFull - list_floor) >> 3;
    span = span < 1ull ? 1ull : span;
    comp_hi = 0xFFFFFFFFFFFFFFFFull;
  }

  while (true) {
    // ---- greedy rounds on the candidates in registers ----
    while (n_sel < limit) {
      u64 best = 0ull;
#pragma unroll
      for (int j = 0; j < kWarpSlots; ++j)
        if (((alive >> j) & 1u) && comp[j] > best) best = comp[j];
      if (__ballot_sync(0xFFFFFFFFu, best != 0ull) == 0u) break;  // nobody alive
      const unsigned hi = __reduce_max_sync(0xFFFFFFFFu, (unsigned)(best >> 32));
      const unsigned lo = __reduce_max_sync(0xFFFFFFFFu, (unsigned)(best >> 32) == hi ? (unsigned)best : 0u);
      const u64 top = ((u64)hi << 32) | (u64)lo;  // composites are unique: exactly one lane owns it
      const int owner = __ffs(__ballot_sync(0xFFFFFFFFu, best == top)) - 1;
      float4 tb = make_float4(0.f, 0.f, 0.f, 0.f);
      float ta = 0.0f;
#pragma unroll
      for (int j = 0; j < kWarpSlots; ++j)
        if (comp[j] == top && ((alive >> j) & 1u)) {
          tb = box[j];
          ta = area[j];
          alive &= ~(1u << j);
        }
      tb.x = __shfl_sync(0xFFFFFFFFu, tb.x, owner);
      tb.y = __shfl_sync(0xFFFFFFFFu, tb.y, owner);
      tb.z = __shfl_sync(0xFFFFFFFFu, tb.z, owner);
      tb.w = __shfl_sync(0xFFFFFFFFu, tb.w, owner);
      ta = __shfl_sync(0xFFFFFFFFu, ta, owner);
      if (lane == n_sel) {
        kbox = tb;
        karea = ta;
        if (fuse_merge) s_picks[warp][n_sel] = top;
        else my_picks[n_sel] = top;
      }
      ++n_sel;
      suppress_by(tb, ta);
    }
    if (n_sel >= limit || comp_hi <= (1ull << 32)) break;  // done, or every candidate has been visited
    // ---- the tier ran dry: next band [comp_lo, comp_hi), from the list while it reaches that far down, else from
    //      the scores of this class ----
    int n_new = 0;
    while (true) {
      const bool from_list = list_ok && n_listed > kWarpCap && comp_hi > list_floor;
      const u64 floor = from_list ? list_floor : (1ull << 32);
      const u64 comp_lo = (comp_hi - floor > span) ? comp_hi - span : floor;
      const unsigned key_lo = (unsigned)(comp_lo >> 32), key_hi32 = (unsigned)((comp_hi - 1ull) >> 32);
      n_new = 0;
      __syncwarp();
      auto offer = [&](const bool in, const u64 cp) {
        const unsigned m = __ballot_sync(0xFFFFFFFFu, in);
        if (in) {
          const int pos = n_new + __popc(m & lt);
          if (pos < kWarpCap) stage[pos] = cp;
        }
        n_new += __popc(m);
      };
      if (from_list) {
        for (int i0 = 0; i0 < n_listed && n_new <= kWarpCap; i0 += 32) {
          const u64 cp = (i0 + lane < n_listed) ? __ldg(lst + i0 + lane) : 0ull;
          offer(cp >= comp_lo && cp < comp_hi, cp);
        }
      } else {
        for (int a0 = 0; a0 < A && n_new <= kWarpCap; a0 += 32 * 16) {  // 16 independent loads in flight per lane
          float sv[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int a = a0 + u * 32 + lane;
            if (LOGITS) {
              const float4 p = (a < A) ? ssdx_softmax4(__ldg(reinterpret_cast<const float4*>(cls_scores - c) + a))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
              sv[u] = (a < A) ? __fmul_rn(c == 0 ? p.x : (c == 1 ? p.y : (c == 2 ? p.z : p.w)), cf) : -INFINITY;
            } else {
              sv[u] = (a < A) ? __fmul_rn(__ldg(cls_scores + (size_t)a * C), cf) : -INFINITY;
            }
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            bool in = false;
            u64 cp = 0ull;
            if (sv[u] > score_thr && a0 + u * 32 + lane < A) {
              const unsigned key = score_to_key(sv[u]);
              if (key >= key_lo && key <= key_hi32) {
                cp = make_comp(key, a0 + u * 32 + lane);
                in = cp >= comp_lo && cp < comp_hi;
              }
            }
            offer(in, cp);
          }
        }
      }
      if (n_new > kWarpCap) {  // over-full: narrow the band and rescan
        span = (span >> 2) > 1ull ? (span >> 2) : 1ull;
        continue;
      }
      comp_hi = comp_lo;
      if (n_new < kWarpRefillTarget / 8) span = (span > 0x1FFFFFFFFFFFFFFFull) ? 0x7FFFFFFFFFFFFFFFull : span << 2;
      else if (n_new < kWarpRefillTarget / 2) span = (span > 0x3FFFFFFFFFFFFFFFull) ? 0x7FFFFFFFFFFFFFFFull : span << 1;
      else if (n_new > kWarpRefillTarget * 2) span = (span >> 1) > 1ull ? (span >> 1) : 1ull;
      if (n_new > 0 || comp_lo <= (1ull << 32)) break;
    }
    __syncwarp();
    alive = 0u;
#pragma unroll
    for (int j = 0; j < kWarpSlots; ++j) comp[j] = (j * 32 + lane < n_new) ? stage[j * 32 + lane] : 0ull;
    load_boxes(n_new);
    // the newcomers against the boxes kept so far
    for (int k = 0; k < n_sel; ++k) {
      float4 kb;
      kb.x = __shfl_sync(0xFFFFFFFFu, kbox.x, k);
      kb.y = __shfl_sync(0xFFFFFFFFu, kbox.y, k);
      kb.z = __shfl_sync(0xFFFFFFFFu, kbox.z, k);
      kb.w = __shfl_sync(0xFFFFFFFFu, kbox.w, k);
      const float ka = __shfl_sync(0xFFFFFFFFu, karea, k);
      suppress_by(kb, ka);
    }
  }
  if (lane == 0 && !fuse_merge) pick_cnt[seg] = n_sel;
  }  // segment
  if (!fuse_merge) return;
  if (lane == 0) s_cnt[warp] = n_sel;  // 0 for the warps beyond the last segment
  __syncthreads();
  const int imgs = kWarpSegs / C;  // images of this CTA: warp w writes the rows of image w
  const int b_img = blockIdx.x * imgs + warp;
  if (warp < imgs && (size_t)b_img * C < (size_t)num_segments)
    merge_image_rows(boxes, &s_picks[warp * C][0], &s_cnt[warp * C], 32, b_img, A, C, limit, K, out, out_valid, out_idx,
                     lane, 32);
}

// ------------------------------------------------------------------------------------------ 1m. merged select
// info per image: [0] listed candidates (may exceed kCap), [1] threshold key (every candidate with key >= it is listed)
template <bool VEC4, bool LOGITS, int NT>  // NT threads per image: 256, or 512 / 1024 when the batch leaves SMs idle
__global__ void __launch_bounds__(NT)
    ssdx_nms_select_all_kernel(const float* __restrict__ scores, const float* __restrict__ cscale, const int A,
                               const int C, const float score_thr, const int all_rank, u64* __restrict__ lists,
                               int* __restrict__ info) {
  constexpr bool logits = LOGITS;
  // LOGITS (VEC4 only): `scores` holds raw logits, the probabilities are ssdx_softmax4 of each row, computed as
  // the rows are read (never written to memory)
  __shared__ unsigned s_max[NT];  // per-thread maximum of the sampled keys
  __shared__ unsigned s_T;
  __shared__ int s_cnt;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const int NS = (A + kSampleStride - 1) / kSampleStride;
  const float* img = scores + (size_t)b * A * C;
  // programmatic dependent launch: the CTAs of fast_all may take the SM slots this grid leaves free right away
  // (they wait in griddepcontrol.wait until this grid has completed and its lists are visible)
  asm volatile("griddepcontrol.launch_dependents;");
  if (tid == 0) s_cnt = 0;

  // ---- sample: every 16th anchor (jittered), all classes; each thread keeps the largest key it sees ----
  {
    unsigned kmax = 0u;  // 0 = no candidate
    for (int k = tid; k < NS; k += NT) {
      int a = k * kSampleStride + ((k * 7) & (kSampleStride - 1));
      a = a < A ? a : A - 1;
      if (VEC4) {
        float4 raw = __ldg(reinterpret_cast<const float4*>(img) + a);
        if (logits) raw = ssdx_softmax4(raw);
        const float4 s = mul4(raw, class_factor4(cscale));
        const float m = fmaxf(fmaxf(s.x, s.y), fmaxf(s.z, s.w));
        if (m > score_thr) kmax = max(kmax, score_to_key(m));
      } else {
        for (int c = 0; c < C; ++c) {
          const float s = __fmul_rn(__ldg(img + (size_t)a * C + c), class_factor(cscale, c));
          if (s > score_thr) kmax = max(kmax, score_to_key(s));
        }
      }
    }
    s_max[tid] = kmax;
  }
  __syncthreads();
  // ---- threshold: the kAllRank-th largest of the per-thread maxima.  At least kAllRank sampled keys are at or
  //      above it (usually just that many: the top keys of a sample sit in different threads), so ~16 * kAllRank
  //      candidates get listed.  One warp bisects over the 256 maxima held in registers. ----
  if (tid < 32) {
    unsigned v[NT / 32];
    int nz = 0;
    unsigned kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll
    for (int r = 0; r < NT / 32; ++r) {
      v[r] = s_max[r * 32 + lane];
      nz += v[r] != 0u;
      kmax = max(kmax, v[r]);
      kmin = min(kmin, v[r] != 0u ? v[r] : 0xFFFFFFFFu);
    }
    nz = __reduce_add_sync(0xFFFFFFFFu, nz);
    kmax = __reduce_max_sync(0xFFFFFFFFu, kmax);
    kmin = __reduce_min_sync(0xFFFFFFFFu, kmin);
    unsigned T = 1u;  // list every candidate
    if ((long long)A * C > kCap && nz >= kAllRank) {  // dense enough to need a threshold at all
      unsigned lo = kmin;       // count(max >= kmin) = nz >= rank
      unsigned hi = kmax + 1u;  // count = 0
      while (hi - lo > 1u) {
        const unsigned mid = lo + ((hi - lo) >> 1);
        int cnt = 0;
#pragma unroll
        for (int r = 0; r < NT / 32; ++r) cnt += v[r] >= mid;
        cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
        if (cnt >= all_rank) {
          lo = mid;
          if (cnt <= all_rank + 3) break;  // any threshold in this band will do
        } else {
          hi = mid;
        }
      }
      T = lo;
      if (lo == kmin && nz <= all_rank + 3) T = 1u;  // every thread maximum is in the band: list every candidate
    }
    if (T == 0x80000000u) T = 0x7FFFFFFFu;  // +0.0 -> -0.0: keeps "s >= t_lo" and "key >= T" the same set
    if (lane == 0) s_T = T;
  }
  __syncthreads();
  const unsigned T = s_T;
  // listed  <=>  s > score_thr and key(s) >= T  <=>  s >= t_lo   (one compare per score)
  const float tf = (T <= 1u) ? -INFINITY : key_to_score(T);
  const float t_lo = (tf > score_thr) ? tf : nextafterf(score_thr, INFINITY);
  // ---- one streaming pass; hits (~1 % of the scores) are only marked, then appended in a compact loop ----
  u64* lst = lists + (size_t)b * kCap;
  if (VEC4) {
    constexpr int kBatch = 8;  // independent 128-bit streaming loads in flight per thread
    const float4 cf4 = class_factor4(cscale);
    for (int a0 = tid; a0 < A; a0 += NT * kBatch) {
      float4 sb[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int a = a0 + u * NT;
        sb[u] = (a < A) ? __ldcs(reinterpret_cast<const float4*>(img) + a)
                        : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      }
      if (logits) {
#pragma unroll
        for (int u = 0; u < kBatch; ++u)
          if (a0 + u * NT < A) sb[u] = ssdx_softmax4(sb[u]);
      }
      unsigned hits = 0u;  // bit 4*u + c
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        sb[u] = mul4(sb[u], cf4);
        hits |= (sb[u].x >= t_lo ? 1u : 0u) << (4 * u);
        hits |= (sb[u].y >= t_lo ? 2u : 0u) << (4 * u);
        hits |= (sb[u].z >= t_lo ? 4u : 0u) << (4 * u);
        hits |= (sb[u].w >= t_lo ? 8u : 0u) << (4 * u);
      }
      while (hits != 0u) {  // the score is re-read (L1/L2 hit) instead of indexing the register batch dynamically
        const int bit = __ffs(hits) - 1;
        hits &= hits - 1u;
        const int a = a0 + (bit >> 2) * NT, c = bit & 3;
        float s;
        if (logits) {
          const float4 p = ssdx_softmax4(__ldg(reinterpret_cast<const float4*>(img) + a));
          s = __fmul_rn(c == 0 ? p.x : (c == 1 ? p.y : (c == 2 ? p.z : p.w)), class_factor(cscale, c));
        } else {
          s = __fmul_rn(__ldg(img + (size_t)a * 4 + c), class_factor(cscale, c));
        }
        const int pos = atomicAdd(&s_cnt, 1);
        if (pos < kCap) lst[pos] = make_comp_all(score_to_key(s), c, a);
      }
    }
  } else {
    const int total = A * C;  // flat coalesced walk
    for (int e = tid; e < total; e += NT) {
      const float s = cscale ? __fmul_rn(__ldcs(img + e), __ldg(cscale + e % C)) : __ldcs(img + e);
      if (s >= t_lo) {
        const int a = e / C;
        const int pos = atomicAdd(&s_cnt, 1);
        if (pos < kCap) lst[pos] = make_comp_all(score_to_key(s), e - a * C, a);
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    info[(size_t)b * kInfo + 0] = s_cnt;
    info[(size_t)b * kInfo + 1] = (int)T;
  }
}

// ------------------------------------------------------------------------------------------ 2m. merged fast path
// One refill scan of fast_all: every score of the image whose composite lies in [comp_lo, comp_hi) is appended to
// s_list (*s_n counts them, also beyond kCap).  Out of line: it runs for the few images whose first list runs dry, and
// kept inline it would sit (twice unrolled) between the greedy pass and the output rows of every image.
template <bool LOGITS, int NT>
__device__ __noinline__ void fast_all_refill_scan(const float* __restrict__ img_scores, const float* __restrict__ cscale,
                                                  const int A, const int C, const float score_thr, const u64 comp_lo,
                                                  const u64 comp_hi, u64* s_list, int* s_n, const int tid) {
  const unsigned key_lo = (unsigned)(comp_lo >> 32), key_hi32 = (unsigned)((comp_hi - 1ull) >> 32);
  auto consider = [&](const float sc, const int a, const int c) {
    if (sc > score_thr) {  // false for the -inf padding
      const unsigned key = score_to_key(sc);
      if (key >= key_lo && key <= key_hi32) {  // cheap 32-bit pre-test, then the exact composite band
        const u64 cp = make_comp_all(key, c, a);
        if (cp >= comp_lo && cp < comp_hi) {
          const int pos = atomicAdd(s_n, 1);
          if (pos < kCap) s_list[pos] = cp;
        }
      }
    }
  };
  if (LOGITS) {  // (C == 4) rows of logits: probabilities recomputed as in the select kernel
    const float4 cf4 = class_factor4(cscale);
    for (int a0 = tid; a0 < A; a0 += NT * 4) {
      float4 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int a = a0 + u * NT;
        r[u] = (a < A) ? __ldg(reinterpret_cast<const float4*>(img_scores) + a) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int a = a0 + u * NT;
        if (a < A) {
          const float4 p = mul4(ssdx_softmax4(r[u]), cf4);
          consider(p.x, a, 0);
          consider(p.y, a, 1);
          consider(p.z, a, 2);
          consider(p.w, a, 3);
        }
      }
    }
  } else {
    const int total = A * C;
    for (int e0 = tid; e0 < total; e0 += NT * 8) {  // 8 independent loads in flight per thread
      float sv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = e0 + u * NT;
        sv[u] = (e < total)
                    ? (cscale ? __fmul_rn(__ldg(img_scores + e), __ldg(cscale + e % C)) : __ldg(img_scores + e))
                    : -INFINITY;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = e0 + u * NT;
        const int a = e / C;
        consider(sv[u], a, e - a * C);
      }
    }
  }
}

// MAXSEL >= K; MINB resident CTAs per SM (7 * 148 >= 1024: one wave at batch 1024); LOGITS: `scores` holds raw logits
// of 4 classes.  ALGO 1 (default): the sorted candidates are taken in chunks of up to 256; a chunk is tested against
// the boxes kept by earlier chunks and against itself as 32 x 32 tiles of a lower-triangular bit matrix (all warps,
// no barrier inside), then one warp resolves the chunk 64 candidates at a time from the bit rows.  ALGO 0
// (ssdx_set_option("nms_path", "blocked")): 64 candidates at a time, each against the kept boxes, two barriers per 64.
template <int MAXSEL, int MINB, bool LOGITS, int ALGO, int NT>
__global__ void __launch_bounds__(NT, MINB)
    ssdx_nms_fast_all_kernel(const ssdx_box_src boxes, const float* __restrict__ scores,
                             const float* __restrict__ cscale, const u64* __restrict__ lists, const int* __restrict__ info, const int A, const int C,
                             const int K, const float iou_thr, const float score_thr, float* __restrict__ out,
                             int* __restrict__ out_valid, int* __restrict__ out_idx) {
  static_assert(ALGO == 1 || NT == kFastThreads, "the blocked form runs 256 threads");
  __shared__ __align__(16) u64 s_list[kCap];
  __shared__ __align__(16) u64 s_tmp[kCap];  // second exchange buffer of the sort; the candidate arrays (float4 first) alias it
  __shared__ float4 s_sbox[MAXSEL];
  __shared__ float s_sarea[MAXSEL];
  __shared__ u64 s_pick[MAXSEL];
  __shared__ unsigned char s_scls[MAXSEL];
  __shared__ HBox s_shbox[MAXSEL];  // prefilter copies of the kept boxes
  __shared__ u64 s_row[ALGO == 0 ? kBlk : 1];  // per candidate of the block: earlier candidates (j < i) that suppress it
  __shared__ int s_sup[ALGO == 0 ? kBlk : 1];
  // ALGO 1: tile (g, jb), jb <= g, at index g (g + 1) / 2 + jb: bit jj of word `lane` = candidate 32 jb + jj of the
  // chunk suppresses candidate 32 g + lane (same class, IoU > threshold, earlier in the order)
  __shared__ unsigned s_tile[ALGO == 1 ? kChunkTiles * 32 : 1];
  __shared__ unsigned char s_supb[ALGO == 1 ? kFastThreads : 1];  // suppressed by a box kept in an earlier chunk / beyond m
  __shared__ unsigned s_keptw[8];  // ALGO 1: kept candidates of the chunk, one word per row group ...
  __shared__ int s_kbase[8];       // ... and the slot of the first kept box of each group
  __shared__ int s_nsel, s_n;
  __shared__ unsigned s_mm[2];  // rank_sort_512: smallest / largest score key of the tier
  float4* s_cbox = reinterpret_cast<float4*>(s_tmp);                              // [kFastThreads]
  float* s_carea = reinterpret_cast<float*>(s_cbox + kFastThreads);               // [kFastThreads]
  unsigned char* s_ccls = reinterpret_cast<unsigned char*>(s_carea + kFastThreads);  // [kFastThreads]
  HBox* s_chbox = reinterpret_cast<HBox*>(s_ccls + kFastThreads);                    // [kFastThreads]
  HPair* s_cpair = reinterpret_cast<HPair*>(s_chbox);  // ALGO 1: the same bytes as column pairs [kFastThreads / 2]
  static_assert(((sizeof(float4) + sizeof(float) + 1) * kFastThreads) % 16 == 0, "s_cpair alignment");
  // ALGO 1: how the raw box differs from the normalised one (bit 0: y corners swapped, bit 1: x corners swapped,
  // bit 2: NaN or equal corners - reload the raw box for the output row)
  unsigned char* s_cflag = reinterpret_cast<unsigned char*>(s_chbox + kFastThreads);  // [kFastThreads]
  __shared__ unsigned char s_sflag[ALGO == 1 ? MAXSEL : 1];
  static_assert((sizeof(float4) + sizeof(float) + 1 + sizeof(HBox) + 1) * kFastThreads <= sizeof(u64) * kCap, "s_tmp");

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // launched as a programmatic dependent of select_all (same ssdx_nms call): nothing that kernel wrote is read
  // before this point (a no-op for a plain launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  int n = info[(size_t)b * kInfo + 0];
  // Everything at or above comp_hi (composite order: score key, class, anchor) has been listed.  Candidates are
  // fetched in composite bands [comp_lo, comp_hi), so even thousands of equal scores are split (by class and anchor)
  // into lists of at most kCap entries: this kernel never hands an image over.
  u64 comp_hi = (u64)(unsigned)info[(size_t)b * kInfo + 1] << 32;
  u64 span = 0ull;
  const bool overflow = n > kCap;  // uniform: the first list does not fit; start over with band refills
  if (overflow) {
    n = 0;
    span = (0xFFFFFFFFFFFFFFFFull - comp_hi) >> 2;  // the overflowing list covered [comp_hi, max]: a quarter of it
    comp_hi = 0xFFFFFFFFFFFFFFFFull;
  }
  const size_t box_row0 = (size_t)b * A;
  const float* img_scores = scores + (size_t)b * A * C;
  const u64* lst = lists + (size_t)b * kCap;
  PHASE_T0();
  for (int i = tid; i < n; i += NT) s_list[i] = lst[i];  // (n = 0 after an overflow: refill right away)
  if (tid == 0) s_nsel = 0;
  __syncthreads();
  PHASE_ADD(10);  // list load

  int n_sel = 0;
  int seen = 0;       // candidates consumed so far (ALGO 1: sizes the chunks after the first from the keep rate)
  float unit = 0.0f;  // ALGO 1: class shift of the prefilter boxes, from the coordinate range of the first chunk
  while (true) {
    if constexpr (ALGO == 1) {
      // ---- a tier above 512 candidates would take the 1024-key sort (more than twice the time of the 512-key one,
      //      and the slowest images set the duration of the kernel): raise the lower end of the tier instead - a
      //      bisection over the score keys for a threshold that keeps 320..512 candidates; what is dropped stays
      //      below comp_hi and is fetched again by a refill in the rare case it is needed ----
      if (n > 512) {  // (uniform)
        unsigned key[kCap / NT];
        unsigned kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll
        for (int r = 0; r < kCap / NT; ++r) {
          const int i = r * NT + tid;
          key[r] = (i < n) ? (unsigned)(s_list[i] >> 32) : 0u;  // (listed keys are >= 1)
          if (i < n) kmin = min(kmin, key[r]), kmax = max(kmax, key[r]);
        }
        kmin = __reduce_min_sync(0xFFFFFFFFu, kmin);
        kmax = __reduce_max_sync(0xFFFFFFFFu, kmax);
        if (tid == 0) s_n = 0, s_nsel = -1;  // (as unsigned: the largest key; s_nsel is rewritten before it is read again)
        __syncthreads();
        if (lane == 0) atomicMax(reinterpret_cast<unsigned*>(&s_n), kmax), atomicMin(reinterpret_cast<unsigned*>(&s_nsel), kmin);
        __syncthreads();
        unsigned lo = (unsigned)s_nsel;      // count(key >= lo) = n > 512
        unsigned hi = (unsigned)s_n + 1u;    // count(key >= hi) = 0  (kmax < 2^32 - 1: NaN scores are not listed)
        unsigned cut = 0u;
        int n_cut = 0;
        while (hi - lo > 1u) {
          const unsigned mid = lo + ((hi - lo) >> 1);
          __syncthreads();
          if (tid == 0) s_n = 0;
          __syncthreads();
          int c = 0;
#pragma unroll
          for (int r = 0; r < kCap / NT; ++r) c += key[r] >= mid;
          c = __reduce_add_sync(0xFFFFFFFFu, c);
          if (lane == 0 && c != 0) atomicAdd(&s_n, c);
          __syncthreads();
          const int cnt = s_n;
          if (cnt > 512) {
            lo = mid;
          } else {
            hi = mid;
            cut = mid;
            n_cut = cnt;
            if (cnt >= 320) break;
          }
        }
        if (n_cut > 0) {  // (uniform) keep the n_cut candidates at or above `cut`
          __syncthreads();
          if (tid == 0) s_n = 0;
          __syncthreads();
#pragma unroll
          for (int r = 0; r < kCap / NT; ++r) {
            if (key[r] >= cut) s_tmp[atomicAdd(&s_n, 1)] = s_list[r * NT + tid];
          }
          __syncthreads();
          for (int i = tid; i < n_cut; i += NT) s_list[i] = s_tmp[i];
          n = n_cut;
          comp_hi = (u64)cut << 32;
          __syncthreads();
        }
      }
    }
    // ---- sort the current tier (merged order) ----
    if (ALGO == 1 && n <= 512) rank_sort_512<NT>(s_list, s_tmp, s_tile, s_mm, n, tid);  // (uniform)
    else fast_sort<(NT >= 512 ? 512 : 256)>(s_list, s_tmp, n, tid);
    PHASE_ADD(11);  // sort
    if (span == 0ull) {  // band that held the lower half of this tier -> width expected to hold the refill target
      const u64 c_mid = (n >= 2) ? s_list[n >> 1] : 0ull;
      const u64 c_bot = (n >= 2) ? s_list[n - 1] : 0ull;
      const u64 half = c_mid - c_bot;  // (sorted descending)
      const u64 per = (n >= 2) ? half / (u64)n : 0ull;
      span = (per == 0ull) ? (1024ull << 32)
                           : (per > (0x7FFFFFFFFFFFFFFFull / (2ull * kAllRefillTarget)) ? 0x7FFFFFFFFFFFFFFFull
                                                                                         : per * (2ull * kAllRefillTarget));
    }

    if constexpr (ALGO == 1) {
      for (int g0 = 0; g0 < n && n_sel < K;) {
        // chunk size: what the keep rate seen so far asks for (+ 1/8 + 8), whole warps, at most one per thread.  The
        // first chunk is kept tight (224 candidates for 200 boxes: 28 tiles instead of 36); a small second chunk is
        // cheaper than eight more tiles for every image
        int want = K - n_sel;
        if (seen > 0) {
          want = (int)(((long long)want * seen + n_sel - 1) / n_sel);  // (n_sel >= 1 once seen > 0)
          want += (want >> 3) + 8;
        } else {
          want += (want >> 4) + 4;  // first chunk: room for 1 suppressed candidate in 17; a later chunk brings the rest
        }
        const int m = min(min(kFastThreads, n - g0), (want + 31) & ~31);
        const int R = (m + 31) >> 5;  // row groups of 32 candidates
        const int n_sel0 = n_sel;
        __syncthreads();  // s_tmp (sort buffer / previous candidate arrays) is free
        float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
        float ai = 0.f;
        int cls = 0;
        HBox hb;
        hb.mn = __float2half2_rn(INFINITY);  // overlaps nothing
        hb.mx = __float2half2_rn(-INFINITY);
        if (tid < m) {  // gather the boxes of the next m sorted candidates
          const u64 cp = s_list[g0 + tid];
          const float4 raw = ssdx_load_box(boxes, box_row0, comp_all_idx(cp));
          const bool fy = raw.x > raw.z, fx = raw.y > raw.w;
          const bool odd = !((raw.x < raw.z) | fy) | !((raw.y < raw.w) | fx);
          s_cflag[tid] = (unsigned char)((fy ? 1 : 0) | (fx ? 2 : 0) | (odd ? 4 : 0));
          bi = normalise_box(raw);
          ai = box_area(bi);
          cls = comp_all_cls(cp);
          s_cbox[tid] = bi;
          s_carea[tid] = ai;
          s_ccls[tid] = (unsigned char)cls;
        }
        if (unit == 0.0f) {  // (uniform) first chunk of the image: coordinate range -> class shift
          const float ext = fmaxf(fmaxf(fabsf(bi.x), fabsf(bi.y)), fmaxf(fabsf(bi.z), fabsf(bi.w)));  // (NaN dropped)
          const unsigned wmax = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(ext));
          if (lane == 0 && tid < kFastThreads) s_tile[warp] = wmax;
          __syncthreads();
          unsigned mx = 0u;
#pragma unroll
          for (int w = 0; w < kFastThreads / 32; ++w) mx = max(mx, s_tile[w]);
          unit = class_unit(__uint_as_float(mx));
        }
        if (tid < m) hb = make_hbox_cls(bi, cls, unit);
        if (tid < kFastThreads) {
          store_hpair(s_cpair, tid, hb);
          s_supb[tid] = (unsigned char)(tid >= m);
        }
        __syncthreads();
        PHASE_ADD(12);  // box gather
        // ---- (a) against the boxes kept by earlier chunks: candidate = tid % (32 R), kept boxes strided over the
        //      256 / (32 R) thread groups, newest first ----
        if (n_sel0 > 0) {
          const int rows = 32 * R, parts = NT / rows;
          const int ci = tid % rows, part = tid / rows;
          if (part < parts && ci < m) {
            const float4 b = s_cbox[ci];
            const float a = s_carea[ci];
            const unsigned c = s_ccls[ci];
            const HBox h = load_hbox(s_cpair, ci);
            bool hit = false;
            if (a > 0.0f) {  // a box without area is never suppressed
              int sidx = n_sel0 - 1 - part;
              for (; sidx >= 3 * parts && !hit; sidx -= 4 * parts) {
                const HBox h0 = s_shbox[sidx], h1 = s_shbox[sidx - parts], h2 = s_shbox[sidx - 2 * parts],
                           h3 = s_shbox[sidx - 3 * parts];
                const bool o0 = hboxes_may_overlap(h, h0), o1 = hboxes_may_overlap(h, h1),
                           o2 = hboxes_may_overlap(h, h2), o3 = hboxes_may_overlap(h, h3);
                if (o0 | o1 | o2 | o3) {
                  hit = (o0 && s_scls[sidx] == c && iou_exceeds(b, a, s_sbox[sidx], s_sarea[sidx], iou_thr)) ||
                        (o1 && s_scls[sidx - parts] == c &&
                         iou_exceeds(b, a, s_sbox[sidx - parts], s_sarea[sidx - parts], iou_thr)) ||
                        (o2 && s_scls[sidx - 2 * parts] == c &&
                         iou_exceeds(b, a, s_sbox[sidx - 2 * parts], s_sarea[sidx - 2 * parts], iou_thr)) ||
                        (o3 && s_scls[sidx - 3 * parts] == c &&
                         iou_exceeds(b, a, s_sbox[sidx - 3 * parts], s_sarea[sidx - 3 * parts], iou_thr));
                }
              }
              for (; sidx >= 0 && !hit; sidx -= parts) {
                if (hboxes_may_overlap(h, s_shbox[sidx]) && s_scls[sidx] == c)
                  hit = iou_exceeds(b, a, s_sbox[sidx], s_sarea[sidx], iou_thr);
              }
            }
            if (hit) s_supb[ci] = 1;  // (several parts may store the same value)
          }
        }
        PHASE_ADD(21);  // kept-box tests
        // ---- (b) the chunk against itself: tiles t = g (g + 1) / 2 + jb of the lower triangle, dealt to the warps ----
        {
          const int ntiles = R * (R + 1) / 2;
          for (int t = warp; t < ntiles; t += NT / 32) {
            const int g = kTileRowGroup[t];
            const int jb = t - g * (g + 1) / 2;
            const int i = 32 * g + lane;
            // the row's own corners, each in both halves (rows and columns beyond m overlap nothing)
            const HPair mine = s_cpair[g * 16 + (lane & 15)];
            const bool up = lane >= 16;
            const __half2 r_ymn = up ? __high2half2(mine.ymn) : __low2half2(mine.ymn);
            const __half2 r_ymx = up ? __high2half2(mine.ymx) : __low2half2(mine.ymx);
            const __half2 r_xmn = up ? __high2half2(mine.xmn) : __low2half2(mine.xmn);
            const __half2 r_xmx = up ? __high2half2(mine.xmx) : __low2half2(mine.xmx);
            const HPair* col = s_cpair + jb * 16;
            unsigned ov = 0u;
#pragma unroll
            for (int p = 0; p < 16; ++p) {  // columns 32 jb + p (bit p) and 32 jb + 16 + p (bit 16 + p)
              const HPair c = col[p];
              const unsigned all = __hgt2_mask(r_ymx, c.ymn) & __hgt2_mask(c.ymx, r_ymn) & __hgt2_mask(r_xmx, c.xmn) &
                                   __hgt2_mask(c.xmx, r_xmn);
              ov |= all & (0x00010001u << p);
            }
            if (jb == g) ov &= (1u << lane) - 1u;  // earlier candidates only
            unsigned keep = 0u;
            if (ov != 0u) {  // class + exact test for the few overlapping pairs
              const float4 b = s_cbox[i];
              const float a = s_carea[i];
              const unsigned c = s_ccls[i];
              if (a > 0.0f) {
                while (ov) {  // two pairs per round (the second repeats the first when only one is left)
                  const int j0 = __ffs(ov) - 1;
                  ov &= ov - 1u;
                  const int j1 = ov ? __ffs(ov) - 1 : j0;
                  ov &= ov - 1u;
                  const int c0 = 32 * jb + j0, c1 = 32 * jb + j1;
                  const bool h0 = (s_ccls[c0] == c) & iou_exceeds_flat(b, a, s_cbox[c0], s_carea[c0], iou_thr);
                  const bool h1 = (s_ccls[c1] == c) & iou_exceeds_flat(b, a, s_cbox[c1], s_carea[c1], iou_thr);
                  keep |= ((unsigned)h0 << j0) | ((unsigned)h1 << j1);
                }
              }
            }
            s_tile[t * 32 + lane] = keep;
          }
        }
        PHASE_ADD(22);  // tiles (thread 0's share)
        __syncthreads();
        PHASE_ADD(13);  // pair tests: barrier
        // ---- (c) one warp decides the chunk, 64 candidates at a time (the arbiter favours the highest warp id) ----
        if (warp == NT / 32 - 1) {
          const unsigned me = 1u << lane;
          unsigned keptw[8];
#pragma unroll
          for (int w = 0; w < 8; ++w) keptw[w] = 0u;
          int ns = n_sel0;
#pragma unroll
          for (int sb = 0; sb < 4; ++sb) {
            if (64 * sb < m && ns < K) {  // (uniform)
              const int ga = 2 * sb, gb = 2 * sb + 1;
              const int ta = ga * (ga + 1) / 2, tb = gb * (gb + 1) / 2;
              const int ia = 64 * sb + lane, ib = ia + 32;
              const bool has_b = gb < R;
              bool sa = s_supb[ia] != 0, sbb = !has_b || s_supb[ib] != 0;
#pragma unroll
              for (int w = 0; w < 2 * sb; ++w) {  // suppressed by a box kept earlier in this chunk
                sa |= (s_tile[(ta + w) * 32 + lane] & keptw[w]) != 0u;
                if (has_b) sbb |= (s_tile[(tb + w) * 32 + lane] & keptw[w]) != 0u;
              }
              const unsigned a_lo = s_tile[(ta + ga) * 32 + lane];  // (diagonal tile: bits below the lane only)
              const unsigned b_lo = has_b ? s_tile[(tb + ga) * 32 + lane] : 0u;
              const unsigned b_hi = has_b ? s_tile[(tb + gb) * 32 + lane] : 0u;
              unsigned supp_lo = __ballot_sync(0xFFFFFFFFu, sa);
              unsigned supp_hi = __ballot_sync(0xFFFFFFFFu, sbb);
              unsigned kept_lo = 0u, kept_hi = 0u;
              // fix-point: the lowest undecided candidate is decided in every round
              while ((kept_lo | supp_lo) != 0xFFFFFFFFu || (kept_hi | supp_hi) != 0xFFFFFFFFu) {
                const bool und_a = !((kept_lo | supp_lo) & me), und_b = !((kept_hi | supp_hi) & me);
                const bool a_hit = (a_lo & kept_lo) != 0u;
                const bool b_hit = ((b_lo & kept_lo) | (b_hi & kept_hi)) != 0u;
                const bool a_free = (a_lo & ~supp_lo) == 0u;
                const bool b_free = ((b_lo & ~supp_lo) | (b_hi & ~supp_hi)) == 0u;
                const unsigned nk_lo = __ballot_sync(0xFFFFFFFFu, und_a && !a_hit && a_free);
                const unsigned ns_lo = __ballot_sync(0xFFFFFFFFu, und_a && a_hit);
                const unsigned nk_hi = __ballot_sync(0xFFFFFFFFu, und_b && !b_hit && b_free);
                const unsigned ns_hi = __ballot_sync(0xFFFFFFFFu, und_b && b_hit);
                kept_lo |= nk_lo;
                supp_lo |= ns_lo;
                kept_hi |= nk_hi;
                supp_hi |= ns_hi;
              }
              keptw[ga] = kept_lo;
              keptw[gb] = kept_hi;
              if (lane == 0) {  // kept masks and the slots their first boxes take: the copies are made by all warps below
                s_keptw[ga] = kept_lo;
                s_keptw[gb] = kept_hi;
                s_kbase[ga] = ns;
                s_kbase[gb] = ns + __popc(kept_lo);
              }
              ns += min(K - ns, __popc(kept_lo) + __popc(kept_hi));
            } else if (lane == 0) {
              s_keptw[2 * sb] = 0u;
              s_keptw[2 * sb + 1] = 0u;
            }
          }
          if (lane == 0) s_nsel = ns;
        }
        __syncthreads();
        PHASE_ADD(14);  // block resolve
        if (tid < m) {  // every kept candidate moves itself to its slot (slots beyond K: dropped)
          const unsigned kw = s_keptw[tid >> 5];
          if ((kw >> lane) & 1u) {
            const int r = s_kbase[tid >> 5] + __popc(kw & ((1u << lane) - 1u));
            if (r < K) {
              s_sbox[r] = s_cbox[tid];
              s_sarea[r] = s_carea[tid];
              s_scls[r] = s_ccls[tid];
              s_shbox[r] = load_hbox(s_cpair, tid);
              s_sflag[r] = s_cflag[tid];
              s_pick[r] = s_list[g0 + tid];
            }
          }
        }
        n_sel = s_nsel;
        PHASE_ADD(24);  // kept boxes copied
        seen += m;
        g0 += m;
      }
    } else {
      for (int g0 = 0; g0 < n && n_sel < K; g0 += kFastThreads) {
        const int gn = min(kFastThreads, n - g0);
        __syncthreads();  // s_tmp (sort buffer / previous candidate arrays) is free
        if (tid < gn) {   // gather the boxes of the next 256 sorted candidates
          const u64 cp = s_list[g0 + tid];
          const float4 nb = normalise_box(ssdx_load_box(boxes, box_row0, comp_all_idx(cp)));
          s_cbox[tid] = nb;
          s_carea[tid] = box_area(nb);
          s_ccls[tid] = (unsigned char)comp_all_cls(cp);
          s_chbox[tid] = make_hbox(nb);
        }
        __syncthreads();
        PHASE_ADD(12);  // box gather
        for (int base = 0; base < gn && n_sel < K; base += kBlk) {
          const int m = min(kBlk, gn - base);
          const int ci = tid >> 2, part = tid & 3;
          {
            bool hit = false;
            float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
            float ai = 0.f;
            unsigned cls = 0u;
            HBox hi_box = make_hbox(bi);
            if (ci < m) {
              bi = s_cbox[base + ci];
              ai = s_carea[base + ci];
              cls = s_ccls[base + ci];
              hi_box = s_chbox[base + ci];
            }
            if (ai > 0.0f) {  // a box without area is never suppressed
              // kept boxes of every class, newest first; the class test and the exact IoU only for overlapping pairs
              int s = n_sel - 1 - part;
              for (; s >= 12 && !hit; s -= 16) {
                const HBox h0 = s_shbox[s], h1 = s_shbox[s - 4], h2 = s_shbox[s - 8], h3 = s_shbox[s - 12];
                const bool o0 = hboxes_may_overlap(hi_box, h0), o1 = hboxes_may_overlap(hi_box, h1),
                           o2 = hboxes_may_overlap(hi_box, h2), o3 = hboxes_may_overlap(hi_box, h3);
                if (o0 | o1 | o2 | o3) {
                  hit = (o0 && s_scls[s] == cls && iou_exceeds(bi, ai, s_sbox[s], s_sarea[s], iou_thr)) ||
                        (o1 && s_scls[s - 4] == cls && iou_exceeds(bi, ai, s_sbox[s - 4], s_sarea[s - 4], iou_thr)) ||
                        (o2 && s_scls[s - 8] == cls && iou_exceeds(bi, ai, s_sbox[s - 8], s_sarea[s - 8], iou_thr)) ||
                        (o3 && s_scls[s - 12] == cls && iou_exceeds(bi, ai, s_sbox[s - 12], s_sarea[s - 12], iou_thr));
                }
              }
              for (; s >= 0 && !hit; s -= 4) {
                if (hboxes_may_overlap(hi_box, s_shbox[s]) && s_scls[s] == cls)
                  hit = iou_exceeds(bi, ai, s_sbox[s], s_sarea[s], iou_thr);
              }
            }
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, hit);
            if (part == 0) s_sup[ci] = (ci >= m) || (((bal >> (lane & 28)) & 0xFu) != 0u);
            // earlier candidates j < ci of the same class that suppress ci; this thread tests j = 4*jj + part
            unsigned lo = 0u, hi = 0u;
            if (ai > 0.0f && part < ci) {
              unsigned ov = 0u;
              const int nj = (ci - part + 3) >> 2;  // number of jj with 4*jj + part < ci
              const int nj4 = (nj + 3) & ~3;        // stale slots beyond m / ci are masked below
              for (int j0 = 0; j0 < nj4; j0 += 4) {
  #pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                  ov |= (unsigned)hboxes_may_overlap(hi_box, s_chbox[base + 4 * (j0 + jj) + part]) << (j0 + jj);
              }
              if (nj < 16) ov &= (1u << nj) - 1u;
              while (ov) {  // class + exact test for the few overlapping pairs
                const int jj = __ffs(ov) - 1;
                ov &= ov - 1u;
                const int j = 4 * jj + part;
                if (s_ccls[base + j] == cls && iou_exceeds(bi, ai, s_cbox[base + j], s_carea[base + j], iou_thr)) {
                  if (j < 32) lo |= 1u << j; else hi |= 1u << (j - 32);
                }
              }
            }
            lo |= __shfl_xor_sync(0xFFFFFFFFu, lo, 1);
            hi |= __shfl_xor_sync(0xFFFFFFFFu, hi, 1);
            lo |= __shfl_xor_sync(0xFFFFFFFFu, lo, 2);
            hi |= __shfl_xor_sync(0xFFFFFFFFu, hi, 2);
            if (part == 0) s_row[ci] = ((u64)hi << 32) | (u64)lo;
          }
          __syncthreads();
          PHASE_ADD(13);  // pair tests
          if (warp == kFastThreads / 32 - 1) {  // the arbiter favours the highest warp id: keep the serial part there
            const unsigned lt = (1u << lane) - 1u, me = 1u << lane;
            const u64 ra = s_row[lane], rb = s_row[lane + 32];
            const unsigned a_lo = (unsigned)ra & lt;
            const unsigned b_lo = (unsigned)rb, b_hi = (unsigned)(rb >> 32) & lt;
            unsigned supp_lo = __ballot_sync(0xFFFFFFFFu, s_sup[lane] != 0);
            unsigned supp_hi = __ballot_sync(0xFFFFFFFFu, s_sup[lane + 32] != 0);
            unsigned kept_lo = 0u, kept_hi = 0u;
            // fix-point: the lowest undecided candidate is decided in every round
            while ((kept_lo | supp_lo) != 0xFFFFFFFFu || (kept_hi | supp_hi) != 0xFFFFFFFFu) {
              const bool und_a = !((kept_lo | supp_lo) & me), und_b = !((kept_hi | supp_hi) & me);
              const bool a_hit = (a_lo & kept_lo) != 0u;
              const bool b_hit = ((b_lo & kept_lo) | (b_hi & kept_hi)) != 0u;
              const bool a_free = (a_lo & ~supp_lo) == 0u;
              const bool b_free = ((b_lo & ~supp_lo) | (b_hi & ~supp_hi)) == 0u;
              const unsigned nk_lo = __ballot_sync(0xFFFFFFFFu, und_a && !a_hit && a_free);
              const unsigned ns_lo = __ballot_sync(0xFFFFFFFFu, und_a && a_hit);
              const unsigned nk_hi = __ballot_sync(0xFFFFFFFFu, und_b && !b_hit && b_free);
              const unsigned ns_hi = __ballot_sync(0xFFFFFFFFu, und_b && b_hit);
              kept_lo |= nk_lo;
              supp_lo |= ns_lo;
              kept_hi |= nk_hi;
              supp_hi |= ns_hi;
            }
            const int room = K - n_sel;
            const int n_lo = __popc(kept_lo);
            if (kept_lo & me) {
              const int r = __popc(kept_lo & lt);
              if (r < room) {
                s_sbox[n_sel + r] = s_cbox[base + lane];
                s_sarea[n_sel + r] = s_carea[base + lane];
                s_scls[n_sel + r] = s_ccls[base + lane];
                s_shbox[n_sel + r] = s_chbox[base + lane];
                s_pick[n_sel + r] = s_list[g0 + base + lane];
              }
            }
            if (kept_hi & me) {
              const int r = n_lo + __popc(kept_hi & lt);
              if (r < room) {
                s_sbox[n_sel + r] = s_cbox[base + lane + 32];
                s_sarea[n_sel + r] = s_carea[base + lane + 32];
                s_scls[n_sel + r] = s_ccls[base + lane + 32];
                s_shbox[n_sel + r] = s_chbox[base + lane + 32];
                s_pick[n_sel + r] = s_list[g0 + base + lane + 32];
              }
            }
            if (lane == 0) s_nsel = n_sel + min(room, n_lo + __popc(kept_hi));
          }
          __syncthreads();
          PHASE_ADD(14);  // block resolve
          n_sel = s_nsel;
        }
      }
    }
    if (n_sel >= K || comp_hi <= (1ull << 32)) break;  // done, or every candidate has been visited
    // ---- the tier ran dry: fetch the next composite band [comp_lo, comp_hi) of all classes from the scores.  The
    //      band width adapts to the density seen; an over-full band is narrowed and rescanned (down to one composite,
    //      i.e. one candidate, so this always makes progress). ----
    bool got = false;
    while (!got) {
      const u64 comp_lo = (comp_hi - (1ull << 32) > span) ? comp_hi - span : (1ull << 32);
      __syncthreads();
      if (tid == 0) s_n = 0;
      __syncthreads();
      fast_all_refill_scan<LOGITS, NT>(img_scores, cscale, A, C, score_thr, comp_lo, comp_hi, s_list, &s_n, tid);
      __syncthreads();
      const int n_new = s_n;
      if (n_new > kCap) {
        span = (span >> 2) > 1ull ? (span >> 2) : 1ull;
      } else {
        n = n_new;
        comp_hi = comp_lo;
        if (n_new < kAllRefillTarget / 8) span = (span > 0x1FFFFFFFFFFFFFFFull) ? 0x7FFFFFFFFFFFFFFFull : span << 2;
        else if (n_new < kAllRefillTarget / 2) span = (span > 0x3FFFFFFFFFFFFFFFull) ? 0x7FFFFFFFFFFFFFFFull : span << 1;
        else if (n_new > kAllRefillTarget * 2) span = (span >> 1) > 1ull ? (span >> 1) : 1ull;
        got = n_new > 0 || comp_lo <= (1ull << 32);
      }
    }
    __syncthreads();
    PHASE_ADD(15);  // refill scans
  }
  PHASE_ADD(23);  // (loop exit: up to the barrier)
  __syncthreads();
  PHASE_ADD(19);  // (loop exit: barrier)
  // ---- rows, already in merge order: [label, prob, xmin, ymin, xmax, ymax] (layers.py:155-162), zero padding ----
  float* o = out + (size_t)b * K * 6;
  if constexpr (ALGO == 1) {
    if ((size_t)K * 6 * sizeof(float) <= sizeof(s_tmp)) {  // (uniform) rows staged in shared memory, stored coalesced
      float* s_rows = reinterpret_cast<float*>(s_tmp);
      for (int r = tid; r < n_sel; r += NT) {
        const u64 cp = s_pick[r];
        const int idx = comp_all_idx(cp);
        const unsigned fl = s_sflag[r];
        const float4 nb = s_sbox[r];
        float4 bx = make_float4((fl & 1u) ? nb.z : nb.x, (fl & 2u) ? nb.w : nb.y, (fl & 1u) ? nb.x : nb.z,
                                (fl & 2u) ? nb.y : nb.w);  // the raw (ymin, xmin, ymax, xmax) back from the normalised box
        if (fl & 4u) bx = ssdx_load_box(boxes, box_row0, idx);
        float* row = s_rows + r * 6;
        row[0] = (float)comp_all_cls(cp);
        row[1] = key_to_score((unsigned)(cp >> 32));
        row[2] = bx.y;  // gather [1,0,3,2], layers.py:155
        row[3] = bx.x;
        row[4] = bx.w;
        row[5] = bx.z;
        if (out_idx) out_idx[(size_t)b * K + r] = idx;
      }
      for (int e = n_sel * 6 + tid; e < K * 6; e += NT) s_rows[e] = 0.0f;
      __syncthreads();
      PHASE_ADD(20);  // rows staged
      if ((reinterpret_cast<uintptr_t>(o) & 7) == 0) {  // (K * 6 floats per image: always a whole number of pairs)
        for (int e = tid; e < K * 3; e += NT)
          reinterpret_cast<float2*>(o)[e] = reinterpret_cast<const float2*>(s_rows)[e];
      } else {
        for (int e = tid; e < K * 6; e += NT) o[e] = s_rows[e];
      }
      if (out_idx)
        for (int e = n_sel + tid; e < K; e += NT) out_idx[(size_t)b * K + e] = -1;
      if (out_valid && tid == 0) out_valid[b] = n_sel;
      PHASE_ADD(16);  // rows
#ifdef SSDX_PHASE_TIMERS
      if (tid == 0) atomicAdd(&g_phase_cycles[18], 1ull);
#endif
      return;
    }
  }
  for (int r = tid; r < n_sel; r += NT) {
    const u64 cp = s_pick[r];
    const int idx = comp_all_idx(cp);
    const float4 bx = ssdx_load_box(boxes, box_row0, idx);  // raw (ymin, xmin, ymax, xmax)
    float* row = o + (size_t)r * 6;
    row[0] = (float)comp_all_cls(cp);
    row[1] = key_to_score((unsigned)(cp >> 32));
    row[2] = bx.y;  // gather [1,0,3,2], layers.py:155
    row[3] = bx.x;
    row[4] = bx.w;
    row[5] = bx.z;
    if (out_idx) out_idx[(size_t)b * K + r] = idx;
  }
  for (int e = n_sel * 6 + tid; e < K * 6; e += NT) o[e] = 0.0f;
  if (out_idx)
    for (int e = n_sel + tid; e < K; e += NT) out_idx[(size_t)b * K + e] = -1;
  if (out_valid && tid == 0) out_valid[b] = n_sel;
  PHASE_ADD(16);  // rows
#ifdef SSDX_PHASE_TIMERS
  if (tid == 0) atomicAdd(&g_phase_cycles[18], 1ull);
#endif
}

// ------------------------------------------------------------------------------------------ 3. generic path
struct GenSmem {
  unsigned* keys;  // [A]      0 = not a (remaining) candidate
  u64* list;       // [kChunk] sorted composite keys of the current chunk
  float4* cbox;    // [kChunk] normalised boxes of the chunk
  float* carea;    // [kChunk]
  float4* sbox;    // [limit]  normalised boxes kept so far
  float* sarea;    // [limit]
  u64* rowmask;    // [kBlk]   suppression bit-matrix of the current block
  int* sup;        // [kBlk]   suppressed by an already kept box
  int* ctr;        // [8]      0..2 rotating counters, 3 list_n, 4 n_sel, 5 candidates, 6 min key, 7 max key
};

__host__ __device__ inline size_t gen_smem_bytes(int A, int limit) {
  size_t s = 0;
  s += (size_t)((A + 3) / 4 * 4) * 4;        // keys
  s += (size_t)kChunk * 8;                   // list
  s += (size_t)kChunk * 16;                  // cbox
  s += (size_t)kChunk * 4;                   // carea
  s += (size_t)((limit + 3) / 4 * 4) * 16;   // sbox
  s += (size_t)((limit + 3) / 4 * 4) * 4;    // sarea
  s += (size_t)kBlk * 8;                     // rowmask
  s += (size_t)kBlk * 4;                     // sup
  s += 64;                                   // counters
  return s;
}

template <bool LOGITS>  // LOGITS: `scores` holds raw logits of 4 classes
__global__ void __launch_bounds__(kGenThreads)
    ssdx_nms_generic_kernel(const ssdx_box_src boxes, const float* __restrict__ scores,
                            const float* __restrict__ cscale, const int A, const int C, const int limit, const float iou_thr, const float score_thr,
                            u64* __restrict__ picks, int* __restrict__ pick_cnt, const int* __restrict__ flags,
                            const int num_segments) {
  // persistent CTAs over the worklist of segments the fast path handed over (flags[0] = count, then ids);
  // without a worklist (flags == nullptr) every segment of the grid-stride range is processed
  const int n_work = (flags != nullptr) ? flags[0] : num_segments;
  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
  const int seg_id = (flags != nullptr) ? flags[1 + work] : work;
  const int b = seg_id / C, c = seg_id - b * C;
  __syncthreads();  // shared memory is reused across worklist items
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GenSmem sm;
  {
    unsigned char* p = smem_raw;
    const int A4 = (A + 3) / 4 * 4, L4 = (limit + 3) / 4 * 4;
    sm.cbox = (float4*)p;   p += (size_t)kChunk * 16;
    sm.sbox = (float4*)p;   p += (size_t)L4 * 16;
    sm.list = (u64*)p;      p += (size_t)kChunk * 8;
    sm.rowmask = (u64*)p;   p += (size_t)kBlk * 8;
    sm.keys = (unsigned*)p; p += (size_t)A4 * 4;
    sm.carea = (float*)p;   p += (size_t)kChunk * 4;
    sm.sarea = (float*)p;   p += (size_t)L4 * 4;
    sm.sup = (int*)p;       p += (size_t)kBlk * 4;
    sm.ctr = (int*)p;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t box_row0 = (size_t)b * A;
  const float* img_scores = scores + (size_t)b * A * C + c;
  const float cf = class_factor(cscale, c);
  u64* my_picks = picks + ((size_t)b * C + c) * limit;

  if (tid < 8) sm.ctr[tid] = (tid == 6) ? -1 : 0;  // ctr[6] = min key (unsigned), ctr[7] = max key
  __syncthreads();

  // ---- A. score filter -> keys ----
  {
    int n = 0;
    unsigned kmin = 0xFFFFFFFFu, kmax = 0u;
    for (int i = tid; i < A; i += kGenThreads) {
      const float s = class_score<LOGITS>(img_scores, i, C, c, cf);
      unsigned k = 0u;
      if (s > score_thr) {
        k = score_to_key(s);
        ++n;
        kmin = min(kmin, k);
        kmax = max(kmax, k);
      }
      sm.keys[i] = k;
    }
    n = __reduce_add_sync(0xFFFFFFFFu, n);
    kmin = __reduce_min_sync(0xFFFFFFFFu, kmin);
    kmax = __reduce_max_sync(0xFFFFFFFFu, kmax);
    if (lane == 0 && n > 0) {
      atomicAdd(&sm.ctr[5], n);
      atomicMin((unsigned*)&sm.ctr[6], kmin);
      atomicMax((unsigned*)&sm.ctr[7], kmax);
    }
  }
  __syncthreads();
  int remaining = sm.ctr[5];
  const unsigned key_lo_all = (unsigned)sm.ctr[6], key_hi_all = (unsigned)sm.ctr[7];
  u64 consumed_from = (((u64)key_hi_all << 32) | 0xFFFFFFFFull) + 1ull;  // every composite >= this is consumed
  int n_sel = 0;
  int it = 0;  // bisection iteration counter (rotating slots 0..2)

  while (remaining > 0 && n_sel < limit) {
    // ---- B. threshold selection ----
    u64 T;
    if (remaining <= kChunk) {
      T = 1ull;
    } else {
      u64 lo = (u64)key_lo_all << 32;  // cnt(lo) = remaining > kChunk
      u64 hi = consumed_from;          // cnt(hi) = 0
      T = 0ull;
      while (true) {
        const u64 mid = lo + ((hi - lo) >> 1);
        int cnt = 0;
        for (int i = tid; i < A; i += kGenThreads) cnt += (make_comp(sm.keys[i], i) >= mid) ? 1 : 0;
        cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
        const int slot = it % 3;
        if (lane == 0 && cnt) atomicAdd(&sm.ctr[slot], cnt);
        __syncthreads();
        cnt = sm.ctr[slot];
        if (tid == 0) sm.ctr[(it + 2) % 3] = 0;
        ++it;
        if (cnt > kChunk) {
          lo = mid;
        } else if (cnt < kChunk / 2) {
          hi = mid;
        } else {
          T = mid;
          break;
        }
      }
    }
    // ---- C. compact + sort the chunk ----
    if (tid == 0) sm.ctr[3] = 0;
    __syncthreads();
    for (int i = tid; i < A; i += kGenThreads) {
      const unsigned k = sm.keys[i];
      if (k != 0u) {
        const u64 cp = make_comp(k, i);
        if (cp >= T) {
          const int pos = atomicAdd(&sm.ctr[3], 1);
          sm.list[pos] = cp;
          sm.keys[i] = 0u;
        }
      }
    }
    __syncthreads();
    const int n = sm.ctr[3];
    remaining -= n;
    consumed_from = T;
    bitonic_sort_desc<kGenThreads>(sm.list, n, tid);
    for (int i = tid; i < n; i += kGenThreads) {
      const float4 nb = normalise_box(ssdx_load_box(boxes, box_row0, comp_idx(sm.list[i])));
      sm.cbox[i] = nb;
      sm.carea[i] = box_area(nb);
    }
    __syncthreads();

    // ---- D. greedy suppression, 64 candidates at a time ----
    for (int base = 0; base < n && n_sel < limit; base += kBlk) {
      const int m = min(kBlk, n - base);
      const int ci = tid >> 3, part = tid & 7;
      {
        bool flag = false;
        float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
        float ai = 0.f;
        if (ci < m) {
          bi = sm.cbox[base + ci];
          ai = sm.carea[base + ci];
          for (int s = part; s < n_sel; s += 8)
            if (iou_exceeds(bi, ai, sm.sbox[s], sm.sarea[s], iou_thr)) {
              flag = true;
              break;
            }
        }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, flag);
        if (part == 0) sm.sup[ci] = (ci >= m) || (((bal >> (lane & 24)) & 0xFFu) != 0u);
        unsigned byte = 0u;
        if (ci < m) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int j = part * 8 + jj;
            if (j > ci && j < m && iou_exceeds(bi, ai, sm.cbox[base + j], sm.carea[base + j], iou_thr))
              byte |= (1u << jj);
          }
        }
        reinterpret_cast<unsigned char*>(sm.rowmask)[ci * 8 + part] = (unsigned char)byte;
      }
      __syncthreads();
      if (warp == 0) {
        const u64 row_lo = sm.rowmask[lane], row_hi = sm.rowmask[lane + 32];
        const unsigned sup_lo = __ballot_sync(0xFFFFFFFFu, sm.sup[lane] != 0);
        const unsigned sup_hi = __ballot_sync(0xFFFFFFFFu, sm.sup[lane + 32] != 0);
        u64 alive = ~(((u64)sup_hi << 32) | (u64)sup_lo);
        u64 kept = 0ull;
        int ns = n_sel;
        while (alive != 0ull && ns < limit) {
          const int i = __ffsll((long long)alive) - 1;
          kept |= (1ull << i);
          ++ns;
          const u64 mine = (i < 32) ? row_lo : row_hi;
          const u64 r = __shfl_sync(0xFFFFFFFFu, mine, i & 31);
          alive &= ~r;
          alive &= ~(1ull << i);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = lane + 32 * h;
          if ((kept >> i) & 1ull) {
            const int pos = n_sel + __popcll(kept & ((1ull << i) - 1ull));
            sm.sbox[pos] = sm.cbox[base + i];
            sm.sarea[pos] = sm.carea[base + i];
            my_picks[pos] = sm.list[base + i];
          }
        }
        if (lane == 0) sm.ctr[4] = ns;
      }
      __syncthreads();
      n_sel = sm.ctr[4];
    }
  }
  if (tid == 0) pick_cnt[(size_t)b * C + c] = n_sel;
  }  // worklist loop
}

// ------------------------------------------------------------------------------------------ 4. merge
__global__ void __launch_bounds__(kMergeThreads)
    ssdx_nms_merge_kernel(const ssdx_box_src boxes, const u64* __restrict__ picks,
                          const int* __restrict__ pick_cnt, const int A, const int C, const int limit, const int K,
                          const int stage_in_smem, float* __restrict__ out, int* __restrict__ out_valid,
                          int* __restrict__ out_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int b = blockIdx.x;  // one CTA per image
  const u64* img_picks = picks + (size_t)b * C * limit;
  const int* cnt = pick_cnt + (size_t)b * C;
  if (stage_in_smem) {  // the binary searches below then run on shared memory instead of L2
    u64* s = (u64*)smem_raw;
    for (int e = tid; e < C * limit; e += kMergeThreads) {
      const int c = e / limit, j = e - c * limit;
      s[e] = (j < cnt[c]) ? img_picks[e] : 0ull;
    }
    __syncthreads();
    img_picks = s;
  }
  merge_image_rows(boxes, img_picks, cnt, limit, b, A, C, limit, K, out, out_valid, out_idx, tid, kMergeThreads);
}

inline int nms_limit(int max_per_class, int A) { return max_per_class < A ? max_per_class : A; }

struct NmsWorkspace {
  size_t picks, pick_cnt, lists, info, flags, total;
};

inline NmsWorkspace nms_workspace(int batch, int A, int C, int max_per_class) {
  NmsWorkspace w;
  const size_t seg = (size_t)batch * C;
  const size_t limit = (size_t)nms_limit(max_per_class, A);
  size_t off = 0;
  w.picks = off;    off += ssdx_align_up(seg * limit * sizeof(u64), 256);
  w.pick_cnt = off; off += ssdx_align_up(seg * sizeof(int), 256);
  w.lists = off;    off += ssdx_align_up(seg * kCap * sizeof(u64), 256);
  w.info = off;     off += ssdx_align_up(seg * kInfo * sizeof(int), 256);
  w.flags = off;    off += ssdx_align_up((seg + 1) * sizeof(int), 256);
  w.total = off;
  return w;
}

}  // namespace

#ifdef SSDX_PHASE_TIMERS
// debug builds only: read and reset the phase counters of ssdx_nms_fast_kernel
extern "C" __attribute__((visibility("default"))) int ssdx_debug_phase_cycles(unsigned long long* out32) {
  unsigned long long zero[32] = {0};
  if (cudaMemcpyFromSymbol(out32, g_phase_cycles, sizeof(zero)) != cudaSuccess) return -2;
  if (cudaMemcpyToSymbol(g_phase_cycles, zero, sizeof(zero)) != cudaSuccess) return -2;
  return 0;
}
#endif

extern "C" size_t ssdx_nms_workspace_bytes(int32_t batch, int32_t num_anchors, int32_t num_classes,
                                           int32_t max_per_class, int32_t max_total) {
  (void)max_total;
  if (batch <= 0 || num_anchors <= 0 || num_classes <= 0 || max_per_class <= 0) return 256;
  return nms_workspace(batch, num_anchors, num_classes, max_per_class).total;
}

extern "C" int ssdx_nms(const float* boxes_yxyx, const float* scores, int32_t batch, int32_t num_anchors,
                        int32_t num_classes, int32_t max_per_class, int32_t max_total, float iou_threshold,
                        float score_threshold, float* out, int32_t* out_valid, int32_t* out_anchor_idx,
                        void* workspace, size_t workspace_bytes, int device, ssdx_stream stream) {
  return ssdx_nms_scaled(boxes_yxyx, scores, nullptr, batch, num_anchors, num_classes, max_per_class, max_total,
                         iou_threshold, score_threshold, out, out_valid, out_anchor_idx, workspace, workspace_bytes,
                         device, stream);
}

// the merged one-pass-per-image pipeline applies (ssdx_set_option("nms_path", generic | perclass) forces the other paths)
static bool nms_merged_applies(const int A, const int C, const int max_per_class, const int max_total) {
  const char* force = ssdx_option_nms_path();
  if (force != nullptr && (force[0] == 'g' || force[0] == 'p')) return false;
  return max_per_class >= max_total && C <= 256 && A <= (1 << 24) && (long long)A * C < (1ll << 31) &&
         max_total <= kFastMaxSel;
}

// raw logits can be consumed directly (soft-max inside every kernel that reads the scores) when they are rows of 4
// aligned classes; otherwise ssdx_nms_logits materialises the probabilities first
static bool nms_logits_fusable(const int A, const int C, const int max_per_class, const int max_total,
                               const float* logits) {
  (void)A;
  (void)max_per_class;
  (void)max_total;
  return C == 4 && (((uintptr_t)logits & 15) == 0);
}

// phases: 1 = candidate selection (reads only the scores), 2 = everything that needs the boxes.  ssdx_nms runs both;
// ssdx_nms_select / ssdx_nms_finish expose them separately so that a caller can run phase 1 on another stream while
// the boxes are still being decoded (in the reference's inference graph the two inputs of the NMS layer are
// independent branches, models.py:402-414).
static int nms_run(const int phases, const int logits, const ssdx_box_src* fused_src, const float* boxes_yxyx,
                   const float* scores, const float* class_scale,
                   int32_t batch, int32_t num_anchors, int32_t num_classes, int32_t max_per_class, int32_t max_total,
                   float iou_threshold, float score_threshold, float* out, int32_t* out_valid,
                   int32_t* out_anchor_idx, void* workspace, size_t workspace_bytes, int device, ssdx_stream stream) {
  const float* cscale = class_scale;
  const bool ph1 = (phases & 1) != 0, ph2 = (phases & 2) != 0;
  SSDX_CHECK_ARG(scores != nullptr && (!ph2 || ((boxes_yxyx != nullptr || fused_src != nullptr) && out != nullptr)),
                 "ssdx_nms: NULL pointer argument");
  SSDX_CHECK_ARG(batch >= 0 && num_anchors > 0 && num_classes > 0, "ssdx_nms: bad shape (B=%d, A=%d, C=%d)", batch,
                 num_anchors, num_classes);
  SSDX_CHECK_ARG(max_per_class > 0 && max_total > 0, "ssdx_nms: max_per_class and max_total must be > 0");
  SSDX_CHECK_ARG(iou_threshold >= 0.0f && iou_threshold <= 1.0f, "iou_threshold must be in [0, 1]");
  SSDX_CHECK_ARG(!ph2 || ((uintptr_t)boxes_yxyx & 15) == 0, "ssdx_nms: boxes must be 16-byte aligned");
  SSDX_CHECK_ARG(batch <= 65535 && num_classes <= 65535 && (long long)batch * num_classes < (1ll << 30),
                 "ssdx_nms: batch and classes must be <= 65535 per call");
  if (batch == 0) return SSDX_OK;
  const int A = num_anchors, C = num_classes;
  const int limit = nms_limit(max_per_class, A);
  const NmsWorkspace w = nms_workspace(batch, A, C, max_per_class);
  if (workspace == nullptr || workspace_bytes < w.total) {
    ssdx_set_error("ssdx_nms: workspace too small (%zu < %zu bytes)", workspace_bytes, w.total);
    return SSDX_ERR_WORKSPACE;
  }
  SSDX_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "ssdx_nms: workspace must be 256-byte aligned");
  const size_t gen_smem = gen_smem_bytes(A, limit);
  SSDX_DEVICE(device);
  cudaStream_t s = (cudaStream_t)stream;
  char* base = (char*)workspace;
  u64* picks = (u64*)(base + w.picks);
  int* pick_cnt = (int*)(base + w.pick_cnt);
  u64* lists = (u64*)(base + w.lists);
  int* info = (int*)(base + w.info);
  int* flags = (int*)(base + w.flags);
  ssdx_box_src bx;
  if (fused_src != nullptr) bx = *fused_src;
  else bx = ssdx_box_src{(const float4*)boxes_yxyx, nullptr, nullptr, 0.f, 0.f, 0.f, 0.f};

  const bool vec4 = (C == 4) && (((uintptr_t)scores & 15) == 0);
  // test hooks: "generic" = exact generic kernel alone, "perclass" = per-class fast pipeline even when the merged
  // one applies.  All paths give identical results.
  const char* force = ssdx_option_nms_path();
  const bool force_generic = force != nullptr && force[0] == 'g';
  const int num_segments = batch * C;
  const size_t merge_smem = (size_t)C * limit * sizeof(u64);
  const int stage = merge_smem <= 96 * 1024;
  if (stage)
    SSDX_CUDA(cudaFuncSetAttribute(ssdx_nms_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)merge_smem));

  // ---- merged pipeline: one list / sort / greedy pass per image ----
  const bool merged = nms_merged_applies(A, C, max_per_class, max_total);
  SSDX_CHECK_ARG(!logits || nms_logits_fusable(A, C, max_per_class, max_total, scores),
                 "ssdx_nms: logits on a path that cannot fuse the soft-max");
  if (merged) {
    if (ph1) {
      // first list sized to max_total (rank (3 K + 63) / 16 of the thread maxima, at most kAllRank = the K = 200 tuning)
      int all_rank = (3 * max_total + 63) / 16;
      all_rank = all_rank < 4 ? 4 : (all_rank > kAllRank ? kAllRank : all_rank);
#define SSDX_SELECT_ALL(NT)                                                                                          \
  do {                                                                                                               \
    if (logits)                                                                                                      \
      ssdx_nms_select_all_kernel<true, true, NT><<<batch, NT, 0, s>>>(scores, cscale, A, C, score_threshold, all_rank, \
                                                                      lists, info);                                  \
    else if (vec4)                                                                                                   \
      ssdx_nms_select_all_kernel<true, false, NT><<<batch, NT, 0, s>>>(scores, cscale, A, C, score_threshold,         \
                                                                       all_rank, lists, info);                       \
    else                                                                                                             \
      ssdx_nms_select_all_kernel<false, false, NT><<<batch, NT, 0, s>>>(scores, cscale, A, C, score_threshold,        \
                                                                        all_rank, lists, info);                      \
  } while (0)
      const int sel_sms = ssdx_sm_count(device);  // few images: wider CTAs keep more loads in flight per image
      const int sel_cta = ssdx_option_nms_cta() ? ssdx_option_nms_cta()
                                                : (batch <= sel_sms ? 1024 : (batch <= 3 * sel_sms ? 512 : kSelThreads));
      if (sel_cta == 1024) SSDX_SELECT_ALL(1024);
      else if (sel_cta == 512) SSDX_SELECT_ALL(512);
      else SSDX_SELECT_ALL(kSelThreads);
#undef SSDX_SELECT_ALL
      SSDX_LAUNCH_CHECK("ssdx_nms_select_all_kernel");
    }
    if (!ph2) return SSDX_OK;
    // select_all of this very call is the stream predecessor: launch as its programmatic dependent
    cudaLaunchAttribute pdl_attr[1];
    pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t fa_cfg = {};
    fa_cfg.gridDim = dim3(batch);
    fa_cfg.dynamicSmemBytes = 0;
    fa_cfg.stream = s;
    fa_cfg.attrs = pdl_attr;
#ifdef SSDX_NO_PDL
    fa_cfg.numAttrs = 0;
#else
    fa_cfg.numAttrs = ph1 ? 1 : 0;
#endif
    const u64* lists_c = lists;
    const int* info_c = info;
    // 7 CTAs of 31 KB per SM only fit with the largest shared-memory carve-out: ask for it (a hint the driver may
    // otherwise not choose on its own)
#ifdef SSDX_NO_CARVEOUT
#define SSDX_FAST_ALL_CARVEOUT(MAXSEL, MINB, LOG, ALGO, NT) do {} while (0)
#else
#define SSDX_FAST_ALL_CARVEOUT(MAXSEL, MINB, LOG, ALGO, NT)                                          \
  SSDX_CUDA(cudaFuncSetAttribute(ssdx_nms_fast_all_kernel<MAXSEL, MINB, LOG, ALGO, NT>,              \
                                 cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))
#endif
#define SSDX_FAST_ALL(MAXSEL, MINB, LOG, ALGO, NT)                                                                  \
  do {                                                                                                              \
    fa_cfg.blockDim = dim3(NT);                                                                                     \
    SSDX_FAST_ALL_CARVEOUT(MAXSEL, MINB, LOG, ALGO, NT);                                                            \
    SSDX_CUDA(cudaLaunchKernelEx(&fa_cfg, ssdx_nms_fast_all_kernel<MAXSEL, MINB, LOG, ALGO, NT>, bx, scores, cscale, \
                                 lists_c, info_c, A, C, max_total, iou_threshold, score_threshold, out, out_valid,  \
                                 out_anchor_idx));                                                                  \
  } while (0)
#define SSDX_FAST_ALL_LOG(MAXSEL, MINB, ALGO, NT)           \
  do {                                                      \
    if (logits) SSDX_FAST_ALL(MAXSEL, MINB, true, ALGO, NT); \
    else SSDX_FAST_ALL(MAXSEL, MINB, false, ALGO, NT);       \
  } while (0)
    const bool blocked = force != nullptr && force[0] == 'b';  // test hook: the 64-candidate blocked form
    // few images: one wide CTA per image (the 32 x 32 tiles of a chunk, the kept-box tests, refill scans and rows
    // spread over 32 or 16 warps; sort and gather stay on the first 8); many images: 7 narrow CTAs per SM
    const int sms = ssdx_sm_count(device);
    // (ssdx_set_option("nms_cta", ...) pins the width: the tests run every width at small batch sizes)
    const int cta = ssdx_option_nms_cta() ? ssdx_option_nms_cta() : (batch <= sms ? 1024 : (batch <= 3 * sms ? 512 : 256));
    if (max_total <= 256) {
      if (blocked) SSDX_FAST_ALL_LOG(256, SSDX_FAST_MINB, 0, kFastThreads);
      else if (cta == 1024) SSDX_FAST_ALL_LOG(256, 1, 1, 1024);
      else if (cta == 512) SSDX_FAST_ALL_LOG(256, 3, 1, 512);
      else SSDX_FAST_ALL_LOG(256, SSDX_FAST_MINB, 1, kFastThreads);
    } else {
      if (blocked) SSDX_FAST_ALL_LOG(kFastMaxSel, 6, 0, kFastThreads);
      else if (cta == 1024) SSDX_FAST_ALL_LOG(kFastMaxSel, 1, 1, 1024);
      else SSDX_FAST_ALL_LOG(kFastMaxSel, 5, 1, kFastThreads);
    }
#undef SSDX_FAST_ALL_LOG
#undef SSDX_FAST_ALL_CARVEOUT
#undef SSDX_FAST_ALL
    SSDX_LAUNCH_CHECK("ssdx_nms_fast_all_kernel");
    return SSDX_OK;  // (no fallback: the band refills of fast_all split even massive score ties)
  }

  // ---- per-class pipeline ----
  // (the generic kernel that stands behind it keeps 4 bytes per anchor in shared memory; the merged pipeline above
  // never hands over to it, so only this pipeline is bounded by that)
  SSDX_CHECK_ARG(gen_smem <= 227 * 1024,
                 "ssdx_nms: num_anchors=%d with max_per_class=%d needs %zu B of shared memory (limit 232448)", A,
                 max_per_class, gen_smem);
  SSDX_CUDA(cudaFuncSetAttribute(logits ? ssdx_nms_generic_kernel<true> : ssdx_nms_generic_kernel<false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gen_smem));
  // 1. select
  const bool use_fast = !force_generic;
  // first tier of a dense class: ~16 * rank candidates, sized to the per-class cap (a tier that runs dry is refilled
  // by the fast kernel at the price of a rescan, so keep a margin), never more than the default 320
  int tier_rank = (3 * limit + 63) / 16;
  tier_rank = tier_rank < 4 ? 4 : (tier_rank > kTierRank ? kTierRank : tier_rank);
  if (use_fast && ph1) {
    dim3 gsel((C + kSelGroup - 1) / kSelGroup, batch);
    if (logits)
      ssdx_nms_select_kernel<true, true><<<gsel, kSelThreads, 0, s>>>(scores, cscale, A, C, score_threshold, tier_rank,
                                                                      lists, info, flags);
    else if (vec4)
      ssdx_nms_select_kernel<true, false><<<gsel, kSelThreads, 0, s>>>(scores, cscale, A, C, score_threshold,
                                                                       tier_rank, lists, info, flags);
    else
      ssdx_nms_select_kernel<false, false><<<gsel, kSelThreads, 0, s>>>(scores, cscale, A, C, score_threshold,
                                                                        tier_rank, lists, info, flags);
    SSDX_LAUNCH_CHECK("ssdx_nms_select_kernel");
  }
  if (!ph2) return SSDX_OK;
  const char* force_path = ssdx_option_nms_path();
  const bool no_warp = force_path != nullptr && force_path[0] == 'p' && force_path[1] == 'c';  // "pcta": test hook
  if (use_fast && limit <= 32 && !no_warp) {
    // 2w. one warp per segment, never hands over: select -> warp -> merge
    const int fuse_merge = (kWarpSegs % C) == 0;  // C = 1, 2, 4, 8: the rows are written by the same kernel
    const int warp_grid = (num_segments + kWarpSegs - 1) / kWarpSegs;
#define SSDX_WARP_LAUNCH(LOG, SLOTS, MINB)                                                                         \
  ssdx_nms_fast_warp_kernel<LOG, SLOTS, MINB><<<warp_grid, kFastThreads, 0, s>>>(                                  \
      bx, scores, cscale, lists, info, A, C, num_segments, limit, iou_threshold, score_threshold, picks, pick_cnt, \
      fuse_merge, max_total, out, out_valid, out_anchor_idx)
    if (limit <= 8) {
      if (logits) SSDX_WARP_LAUNCH(true, 8, 1);
      else SSDX_WARP_LAUNCH(false, 8, 1);
    } else {
      if (logits) SSDX_WARP_LAUNCH(true, 4, 4);
      else SSDX_WARP_LAUNCH(false, 4, 4);
    }
#undef SSDX_WARP_LAUNCH
    SSDX_LAUNCH_CHECK("ssdx_nms_fast_warp_kernel");
    if (fuse_merge) return SSDX_OK;
    ssdx_nms_merge_kernel<<<batch, kMergeThreads, stage ? merge_smem : 0, s>>>(bx, picks, pick_cnt, A, C, limit,
                                                                               max_total, stage, out, out_valid,
                                                                               out_anchor_idx);
    SSDX_LAUNCH_CHECK("ssdx_nms_merge_kernel");
    return SSDX_OK;
  }
  if (use_fast) {
    // 2. fast path
    dim3 gseg(C, batch);
    if (logits)
      ssdx_nms_fast_kernel<true><<<gseg, kFastThreads, 0, s>>>(bx, scores, cscale, lists, info, A, C, limit,
                                                               iou_threshold, score_threshold, picks, pick_cnt, flags);
    else
      ssdx_nms_fast_kernel<false><<<gseg, kFastThreads, 0, s>>>(bx, scores, cscale, lists, info, A, C, limit,
                                                                iou_threshold, score_threshold, picks, pick_cnt, flags);
    SSDX_LAUNCH_CHECK("ssdx_nms_fast_kernel");
  }
  // 3. generic path (flag-gated; handles every segment when the fast path is not applicable)
  const int gen_grid = use_fast ? min(num_segments, 2 * ssdx_sm_count(device)) : num_segments;
  if (logits)
    ssdx_nms_generic_kernel<true><<<gen_grid, kGenThreads, gen_smem, s>>>(bx, scores, cscale, A, C, limit, iou_threshold,
                                                                          score_threshold, picks, pick_cnt,
                                                                          use_fast ? flags : nullptr, num_segments);
  else
    ssdx_nms_generic_kernel<false><<<gen_grid, kGenThreads, gen_smem, s>>>(bx, scores, cscale, A, C, limit, iou_threshold,
                                                                           score_threshold, picks, pick_cnt,
                                                                           use_fast ? flags : nullptr, num_segments);
  SSDX_LAUNCH_CHECK("ssdx_nms_generic_kernel");
  // 4. merge
  ssdx_nms_merge_kernel<<<batch, kMergeThreads, stage ? merge_smem : 0, s>>>(bx, picks, pick_cnt, A, C, limit,
                                                                             max_total, stage, out, out_valid,
                                                                             out_anchor_idx);
  SSDX_LAUNCH_CHECK("ssdx_nms_merge_kernel");
  return SSDX_OK;
}

extern "C" int ssdx_nms_scaled(const float* boxes_yxyx, const float* scores, const float* class_scale, int32_t batch,
                               int32_t num_anchors, int32_t num_classes, int32_t max_per_class, int32_t max_total,
                               float iou_threshold, float score_threshold, float* out, int32_t* out_valid,
                               int32_t* out_anchor_idx, void* workspace, size_t workspace_bytes, int device,
                               ssdx_stream stream) {
  return nms_run(3, 0, nullptr, boxes_yxyx, scores, class_scale, batch, num_anchors, num_classes, max_per_class, max_total,
                 iou_threshold, score_threshold, out, out_valid, out_anchor_idx, workspace, workspace_bytes, device,
                 stream);
}

extern "C" int ssdx_nms_select(const float* scores, const float* class_scale, int32_t batch, int32_t num_anchors,
                               int32_t num_classes, int32_t max_per_class, int32_t max_total, float score_threshold,
                               void* workspace, size_t workspace_bytes, int device, ssdx_stream stream) {
  return nms_run(1, 0, nullptr, nullptr, scores, class_scale, batch, num_anchors, num_classes, max_per_class, max_total, 0.5f,
                 score_threshold, nullptr, nullptr, nullptr, workspace, workspace_bytes, device, stream);
}

extern "C" int ssdx_nms_finish(const float* boxes_yxyx, const float* scores, const float* class_scale, int32_t batch,
                               int32_t num_anchors, int32_t num_classes, int32_t max_per_class, int32_t max_total,
                               float iou_threshold, float score_threshold, float* out, int32_t* out_valid,
                               int32_t* out_anchor_idx, void* workspace, size_t workspace_bytes, int device,
                               ssdx_stream stream) {
  return nms_run(2, 0, nullptr, boxes_yxyx, scores, class_scale, batch, num_anchors, num_classes, max_per_class, max_total,
                 iou_threshold, score_threshold, out, out_valid, out_anchor_idx, workspace, workspace_bytes, device,
                 stream);
}

extern "C" int ssdx_decode_nms(const ssdx_anchors* anchors, const float* offsets, const float* std_host,
                               const float* scores, const float* class_scale, int32_t batch, int32_t num_classes,
                               int32_t max_per_class, int32_t max_total, float iou_threshold, float score_threshold,
                               float* out, int32_t* out_valid, int32_t* out_anchor_idx, void* workspace,
                               size_t workspace_bytes, ssdx_stream stream) {
  SSDX_CHECK_ARG(anchors != nullptr && offsets != nullptr && std_host != nullptr,
                 "ssdx_decode_nms: NULL pointer argument");
  SSDX_CHECK_ARG(((uintptr_t)offsets & 15) == 0, "ssdx_decode_nms: offsets must be 16-byte aligned");
  const ssdx_box_src src{nullptr, (const float4*)offsets, anchors->dec_cent, std_host[0], std_host[1], std_host[2],
                         std_host[3]};
  return nms_run(3, 0, &src, nullptr, scores, class_scale, batch, anchors->A, num_classes, max_per_class, max_total,
                 iou_threshold, score_threshold, out, out_valid, out_anchor_idx, workspace, workspace_bytes,
                 anchors->device, stream);
}

// ---- NMS on raw logits: soft-max fused into the score filter ----
extern "C" size_t ssdx_nms_logits_workspace_bytes(int32_t batch, int32_t num_anchors, int32_t num_classes,
                                                  int32_t max_per_class, int32_t max_total) {
  const size_t base = ssdx_nms_workspace_bytes(batch, num_anchors, num_classes, max_per_class, max_total);
  if (base == 0) return 0;
  // room to materialise the probabilities when the fused path does not apply (class count != 4, unaligned rows)
  return ssdx_align_up(base, 256) + ssdx_align_up((size_t)batch * num_anchors * num_classes * sizeof(float), 256);
}

extern "C" int ssdx_nms_logits(const float* boxes_yxyx, const float* logits, const float* class_scale, int32_t batch,
                               int32_t num_anchors, int32_t num_classes, int32_t max_per_class, int32_t max_total,
                               float iou_threshold, float score_threshold, float* out, int32_t* out_valid,
                               int32_t* out_anchor_idx, void* workspace, size_t workspace_bytes, int device,
                               ssdx_stream stream) {
  SSDX_CHECK_ARG(logits != nullptr, "ssdx_nms_logits: NULL pointer argument");
  SSDX_CHECK_ARG(batch >= 0 && num_anchors > 0 && num_classes > 0, "ssdx_nms_logits: bad shape (B=%d, A=%d, C=%d)", batch,
                 num_anchors, num_classes);
  SSDX_CHECK_ARG(max_per_class > 0 && max_total > 0, "ssdx_nms_logits: max_per_class and max_total must be > 0");
  const size_t need = ssdx_nms_logits_workspace_bytes(batch, num_anchors, num_classes, max_per_class, max_total);
  if (workspace == nullptr || workspace_bytes < need) {
    ssdx_set_error("ssdx_nms_logits: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return SSDX_ERR_WORKSPACE;
  }
  const bool fused = nms_logits_fusable(num_anchors, num_classes, max_per_class, max_total, logits);
  if (fused)  // the probabilities exist only in registers
    return nms_run(3, 1, nullptr, boxes_yxyx, logits, class_scale, batch, num_anchors, num_classes, max_per_class,
                   max_total, iou_threshold, score_threshold, out, out_valid, out_anchor_idx, workspace, workspace_bytes,
                   device, stream);
  const size_t base = ssdx_align_up(
      ssdx_nms_workspace_bytes(batch, num_anchors, num_classes, max_per_class, max_total), 256);
  float* probs = (float*)((char*)workspace + base);
  const int rc = ssdx_softmax_scores(logits, (int64_t)batch * num_anchors, num_classes, nullptr, probs, device, stream);
  if (rc != SSDX_OK) return rc;
  return nms_run(3, 0, nullptr, boxes_yxyx, probs, class_scale, batch, num_anchors, num_classes, max_per_class,
                 max_total, iou_threshold, score_threshold, out, out_valid, out_anchor_idx, workspace, base, device,
                 stream);
}

// ---- the whole inference tail in one call: offsets + logits (+ class scale) -> detections.  Boxes are decoded per
//      candidate (ssdx_decode_nms), probabilities are computed inside the score filter (ssdx_nms_logits): neither the
//      (B,A,4) box tensor nor the (B,A,C) probability tensor is written.  Same workspace as ssdx_nms_logits. ----
extern "C" int ssdx_decode_nms_logits(const ssdx_anchors* anchors, const float* offsets, const float* std_host,
                                      const float* logits, const float* class_scale, int32_t batch, int32_t num_classes,
                                      int32_t max_per_class, int32_t max_total, float iou_threshold,
                                      float score_threshold, float* out, int32_t* out_valid, int32_t* out_anchor_idx,
                                      void* workspace, size_t workspace_bytes, ssdx_stream stream) {
  SSDX_CHECK_ARG(anchors != nullptr && offsets != nullptr && std_host != nullptr && logits != nullptr,
                 "ssdx_decode_nms_logits: NULL pointer argument");
  SSDX_CHECK_ARG(((uintptr_t)offsets & 15) == 0, "ssdx_decode_nms_logits: offsets must be 16-byte aligned");
  SSDX_CHECK_ARG(batch >= 0 && num_classes > 0 && max_per_class > 0 && max_total > 0,
                 "ssdx_decode_nms_logits: bad shape or caps");
  const int A = anchors->A;
  const size_t need = ssdx_nms_logits_workspace_bytes(batch, A, num_classes, max_per_class, max_total);
  if (workspace == nullptr || workspace_bytes < need) {
    ssdx_set_error("ssdx_decode_nms_logits: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return SSDX_ERR_WORKSPACE;
  }
  const ssdx_box_src src{nullptr, (const float4*)offsets, anchors->dec_cent, std_host[0], std_host[1], std_host[2],
                         std_host[3]};
  const bool fused = nms_logits_fusable(A, num_classes, max_per_class, max_total, logits);
  if (fused)
    return nms_run(3, 1, &src, nullptr, logits, class_scale, batch, A, num_classes, max_per_class, max_total,
                   iou_threshold, score_threshold, out, out_valid, out_anchor_idx, workspace, workspace_bytes,
                   anchors->device, stream);
  const size_t base = ssdx_align_up(ssdx_nms_workspace_bytes(batch, A, num_classes, max_per_class, max_total), 256);
  float* probs = (float*)((char*)workspace + base);
  const int rc = ssdx_softmax_scores(logits, (int64_t)batch * A, num_classes, nullptr, probs, anchors->device, stream);
  if (rc != SSDX_OK) return rc;
  return nms_run(3, 0, &src, nullptr, probs, class_scale, batch, A, num_classes, max_per_class, max_total,
                 iou_threshold, score_threshold, out, out_valid, out_anchor_idx, workspace, base, anchors->device,
                 stream);
}
