// Per-class confidence-filtered NMS + cross-class merge
// (reference: layers.py:127-168 -> tf.image.combined_non_max_suppression, q = 1, clip_boxes = False).
//
// Kernel 1, one CTA per (image, class):
//   A. score filter (score > thr, strict) -> order-preserving 32-bit keys in shared memory;
//   B. top-k' selection instead of a full sort: a 64-bit bisection on the composite key
//      (score key << 32 | ~anchor) finds a threshold that leaves between kChunk/2 and kChunk
//      candidates (all composite keys are distinct, so any tie pattern - e.g. thousands of scores equal
//      to 1.0f - terminates and is visited lower anchor first);
//   C. those candidates are compacted and bitonic-sorted in shared memory;
//   D. greedy suppression in blocks of 64 sorted candidates: every candidate of the block is tested against
//      the boxes already kept (8 threads per candidate), a 64x64 suppression bit-matrix of the block is built
//      (one byte per thread), and one warp sweeps it cooperatively (ffs / shuffle / and-not), so the serial
//      part is one step per KEPT box.  Stops at min(max_per_class, A); if the chunk runs dry the CTA goes
//      back to B for the next k' candidates, so arbitrarily deep suppression chains stay exact.
// Kernel 2, one CTA per image: rank-merges the C sorted pick lists (score desc; ties: lower class, lower
//   anchor), keeps max_total rows and writes [label, prob, xmin, ymin, xmax, ymax] (layers.py:155-162),
//   zero padding, valid counts and source anchor indices.
//
// IoU follows TF's IOU(): min/max-normalised corners, area = (y1-y0)*(x1-x0), 0 if either area <= 0,
// inter / ((area_i + area_j) - inter) with IEEE division, suppressed iff IoU > iou_thr (strict).
#include "common.cuh"

namespace {

constexpr int kNmsThreads = 512;
constexpr int kChunk = 512;
constexpr int kBlk = 64;
constexpr int kMergeThreads = 256;

typedef unsigned long long u64;

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

// box normalised as TF does: (y0, x0, y1, x1) with y0<=y1, x0<=x1
__device__ __forceinline__ float4 normalise_box(const float4 b) {
  return make_float4(fminf(b.x, b.z), fminf(b.y, b.w), fmaxf(b.x, b.z), fmaxf(b.y, b.w));
}
__device__ __forceinline__ float box_area(const float4 n) {
  return __fmul_rn(__fsub_rn(n.z, n.x), __fsub_rn(n.w, n.y));
}
__device__ __forceinline__ bool iou_exceeds(const float4 a, const float aa, const float4 b, const float ab,
                                            const float thr) {
  if (!(aa > 0.0f) || !(ab > 0.0f)) return false;  // TF: IoU = 0, and thr >= 0
  const float ih = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
  const float iw = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
  const float inter = __fmul_rn(ih, iw);
  if (!(inter > 0.0f)) return false;  // IoU == 0 is never > thr >= 0
  const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
  return iou > thr;
}

struct NmsSmem {
  unsigned* keys;   // [A]      0 = not a (remaining) candidate
  u64* list;        // [kChunk] sorted composite keys of the current chunk
  float4* cbox;     // [kChunk] normalised boxes of the chunk
  float* carea;     // [kChunk]
  float4* sbox;     // [limit]  normalised boxes kept so far
  float* sarea;     // [limit]
  u64* rowmask;     // [kBlk]   suppression bit-matrix of the current block
  int* sup;         // [kBlk]   suppressed by an already kept box
  int* ctr;         // [8]      0..2 rotating counters, 3 list_n, 4 n_sel, 5 candidates, 6 min key, 7 max key
};

__host__ __device__ inline size_t nms_smem_bytes(int A, int limit) {
  size_t s = 0;
  s += (size_t)((A + 3) / 4 * 4) * 4;  // keys
  s += (size_t)kChunk * 8;             // list
  s += (size_t)kChunk * 16;            // cbox
  s += (size_t)kChunk * 4;             // carea
  s += (size_t)((limit + 3) / 4 * 4) * 16;  // sbox
  s += (size_t)((limit + 3) / 4 * 4) * 4;   // sarea
  s += (size_t)kBlk * 8;               // rowmask
  s += (size_t)kBlk * 4;               // sup
  s += 64;                             // counters
  return s;
}

__global__ void __launch_bounds__(kNmsThreads)
    ssdx_nms_class_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, const int A, const int C,
                          const int limit, const float iou_thr, const float score_thr, u64* __restrict__ picks,
                          int* __restrict__ pick_cnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NmsSmem sm;
  {
    unsigned char* p = smem_raw;
    const int A4 = (A + 3) / 4 * 4, L4 = (limit + 3) / 4 * 4;
    sm.cbox = (float4*)p;  p += (size_t)kChunk * 16;
    sm.sbox = (float4*)p;  p += (size_t)L4 * 16;
    sm.list = (u64*)p;     p += (size_t)kChunk * 8;
    sm.rowmask = (u64*)p;  p += (size_t)kBlk * 8;
    sm.keys = (unsigned*)p; p += (size_t)A4 * 4;
    sm.carea = (float*)p;  p += (size_t)kChunk * 4;
    sm.sarea = (float*)p;  p += (size_t)L4 * 4;
    sm.sup = (int*)p;      p += (size_t)kBlk * 4;
    sm.ctr = (int*)p;
  }
  const int c = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* img_boxes = boxes + (size_t)b * A;
  const float* img_scores = scores + (size_t)b * A * C + c;
  u64* my_picks = picks + ((size_t)b * C + c) * limit;

  if (tid < 8) sm.ctr[tid] = (tid == 6) ? -1 : 0;  // ctr[6] = min key (unsigned), ctr[7] = max key
  __syncthreads();

  // ---- A. score filter -> keys ----
  {
    int n = 0;
    unsigned kmin = 0xFFFFFFFFu, kmax = 0u;
    for (int i = tid; i < A; i += kNmsThreads) {
      const float s = __ldg(img_scores + (size_t)i * C);
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
        for (int i = tid; i < A; i += kNmsThreads) cnt += (make_comp(sm.keys[i], i) >= mid && sm.keys[i] != 0u) ? 1 : 0;
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
    for (int i = tid; i < A; i += kNmsThreads) {
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
    for (int i = n + tid; i < kChunk; i += kNmsThreads) sm.list[i] = 0ull;
    __syncthreads();
    // bitonic sort, descending, kChunk elements, one compare-exchange per thread pair
    for (int k = 2; k <= kChunk; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = tid; t < kChunk; t += kNmsThreads) {
          const int p = t ^ j;
          if (p > t) {
            const u64 x = sm.list[t], y = sm.list[p];
            const bool desc = ((t & k) == 0);
            if (desc ? (x < y) : (x > y)) {
              sm.list[t] = y;
              sm.list[p] = x;
            }
          }
        }
        __syncthreads();
      }
    }
    // boxes of the chunk
    for (int i = tid; i < n; i += kNmsThreads) {
      const float4 nb = normalise_box(__ldg(img_boxes + comp_idx(sm.list[i])));
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
}

// number of entries of a descending list whose score key is > k (strict) or >= k
__device__ __forceinline__ int count_before(const u64* __restrict__ lst, int n, unsigned k, bool inclusive) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const unsigned km = (unsigned)(lst[mid] >> 32);
    const bool before = inclusive ? (km >= k) : (km > k);
    if (before) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kMergeThreads)
    ssdx_nms_merge_kernel(const float4* __restrict__ boxes, const u64* __restrict__ picks,
                          const int* __restrict__ pick_cnt, const int A, const int C, const int limit, const int K,
                          float* __restrict__ out, int* __restrict__ out_valid, int* __restrict__ out_idx) {
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const u64* img_picks = picks + (size_t)b * C * limit;
  const int* cnt = pick_cnt + (size_t)b * C;
  int total = 0;
  for (int c = 0; c < C; ++c) total += cnt[c];
  const int valid = min(total, K);
  float* o = out + (size_t)b * K * 6;
  for (int e = tid; e < C * limit; e += kMergeThreads) {
    const int c = e / limit, j = e - c * limit;
    if (j >= cnt[c]) continue;
    const u64 cp = img_picks[(size_t)c * limit + j];
    const unsigned k = (unsigned)(cp >> 32);
    int rank = j;
    for (int c2 = 0; c2 < C; ++c2)
      if (c2 != c) rank += count_before(img_picks + (size_t)c2 * limit, cnt[c2], k, c2 < c);
    if (rank < K) {
      const int idx = comp_idx(cp);
      const float4 bx = __ldg(boxes + (size_t)b * A + idx);  // raw (ymin, xmin, ymax, xmax)
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
  for (int e = valid * 6 + tid; e < K * 6; e += kMergeThreads) o[e] = 0.0f;
  if (out_idx)
    for (int e = valid + tid; e < K; e += kMergeThreads) out_idx[(size_t)b * K + e] = -1;
  if (out_valid && tid == 0) out_valid[b] = valid;
}

inline int nms_limit(int max_per_class, int A) { return max_per_class < A ? max_per_class : A; }

}  // namespace

extern "C" size_t ssdx_nms_workspace_bytes(int32_t batch, int32_t num_anchors, int32_t num_classes,
                                           int32_t max_per_class, int32_t max_total) {
  (void)max_total;
  if (batch <= 0 || num_anchors <= 0 || num_classes <= 0 || max_per_class <= 0) return 256;
  const size_t limit = (size_t)nms_limit(max_per_class, num_anchors);
  return ssdx_align_up((size_t)batch * num_classes * limit * sizeof(u64), 256) +
         ssdx_align_up((size_t)batch * num_classes * sizeof(int), 256);
}

extern "C" int ssdx_nms(const float* boxes_yxyx, const float* scores, int32_t batch, int32_t num_anchors,
                        int32_t num_classes, int32_t max_per_class, int32_t max_total, float iou_threshold,
                        float score_threshold, float* out, int32_t* out_valid, int32_t* out_anchor_idx,
                        void* workspace, size_t workspace_bytes, int device, ssdx_stream stream) {
  SSDX_CHECK_ARG(boxes_yxyx != nullptr && scores != nullptr && out != nullptr, "ssdx_nms: NULL pointer argument");
  SSDX_CHECK_ARG(batch >= 0 && num_anchors > 0 && num_classes > 0, "ssdx_nms: bad shape (B=%d, A=%d, C=%d)", batch,
                 num_anchors, num_classes);
  SSDX_CHECK_ARG(max_per_class > 0 && max_total > 0, "ssdx_nms: max_per_class and max_total must be > 0");
  SSDX_CHECK_ARG(iou_threshold >= 0.0f && iou_threshold <= 1.0f, "iou_threshold must be in [0, 1]");
  SSDX_CHECK_ARG(((uintptr_t)boxes_yxyx & 15) == 0, "ssdx_nms: boxes must be 16-byte aligned");
  SSDX_CHECK_ARG(batch <= 65535 && num_classes <= 65535, "ssdx_nms: batch and classes must be <= 65535 per call");
  if (batch == 0) return SSDX_OK;
  const int limit = nms_limit(max_per_class, num_anchors);
  const size_t need = ssdx_nms_workspace_bytes(batch, num_anchors, num_classes, max_per_class, max_total);
  if (workspace == nullptr || workspace_bytes < need) {
    ssdx_set_error("ssdx_nms: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return SSDX_ERR_WORKSPACE;
  }
  SSDX_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "ssdx_nms: workspace must be 256-byte aligned");
  const size_t smem = nms_smem_bytes(num_anchors, limit);
  SSDX_CHECK_ARG(smem <= 227 * 1024,
                 "ssdx_nms: num_anchors=%d with max_per_class=%d needs %zu B of shared memory (limit 232448)",
                 num_anchors, max_per_class, smem);
  SSDX_CUDA(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  u64* picks = (u64*)workspace;
  int* pick_cnt = (int*)((char*)workspace + ssdx_align_up((size_t)batch * num_classes * limit * sizeof(u64), 256));
  SSDX_CUDA(cudaFuncSetAttribute(ssdx_nms_class_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(num_classes, batch);
  ssdx_nms_class_kernel<<<grid, kNmsThreads, smem, s>>>((const float4*)boxes_yxyx, scores, num_anchors, num_classes,
                                                        limit, iou_threshold, score_threshold, picks, pick_cnt);
  SSDX_LAUNCH_CHECK("ssdx_nms_class_kernel");
  ssdx_nms_merge_kernel<<<batch, kMergeThreads, 0, s>>>((const float4*)boxes_yxyx, picks, pick_cnt, num_anchors,
                                                        num_classes, limit, max_total, out, out_valid,
                                                        out_anchor_idx);
  SSDX_LAUNCH_CHECK("ssdx_nms_merge_kernel");
  return SSDX_OK;
}
