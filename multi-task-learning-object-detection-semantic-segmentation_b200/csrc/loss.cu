// Training losses that consume the encoded targets (reference: losses.py:5-50 localization_loss, :52-172
// confidence_loss), SURVEY 8(f) row 3.  Values per image, plus (optionally) the gradient with respect to the
// predictions that TensorFlow's autodiff derives from the same statements, so that a training loop can take both from
// one pass over the tensors.
//
//   localization  smooth-L1 over the anchors whose target row is not background (sum |y_true| > 0), divided by their
//                 count (at least 1).  One CTA per image: 32 B per anchor in, 16 B out when the gradient is wanted.
//   confidence    soft-max cross-entropy with hard-negative mining: the k = min(3 * positives, negatives) background
//                 anchors with the largest loss are kept, k and the top-k taken over the WHOLE batch (losses.py:117,
//                 :134-138), equal losses lower flat index first (tf.math.top_k).  The k-th largest loss is found by
//                 a radix select (11 + 11 + 10 bits) over order-preserving keys of the per-anchor background losses:
//                 the first histogram is built while the losses are computed, the two others re-read the 4 B keys
//                 (L2-resident: 39 MB for 1024 x 9600).  Ties at the threshold are resolved exactly (an exclusive
//                 scan of the per-image tie counts gives every image its share, taken from its lowest anchors).
//
// Sums are fp32 in a fixed order (deterministic run to run); the reference's reduce_sum order is unspecified, parity is
// 1e-5 relative.  The kept set is index work and exact given the per-anchor losses (logf here vs the reference's log:
// <= 2 ulp, so only anchors whose loss is within a few ulp of the threshold can differ from a CPU run).
#include <cub/block/block_scan.cuh>

#include "common.cuh"

namespace {

constexpr int kLossThreads = 512;
constexpr int kBins0 = 2048, kBins1 = 2048, kBins2 = 1024;  // key bits 31..21, 20..10, 9..0
constexpr int kHistThreads = 512;

// order-preserving map float -> u32 (-0 and +0 share a key, as they compare equal in the reference's sort)
__device__ __forceinline__ unsigned loss_key(float f) {
  const unsigned u = __float_as_uint(__fadd_rn(f, 0.0f));
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_loss(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

__device__ __forceinline__ float block_sum(float v, float* s_buf, const int tid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  __syncthreads();  // s_buf may still be read from a previous call
  if ((tid & 31) == 0) s_buf[tid >> 5] = v;
  __syncthreads();
  float tot = 0.0f;
  for (int w = 0; w < kLossThreads / 32; ++w) tot = __fadd_rn(tot, s_buf[w]);  // every thread, same order
  return tot;
}

// ------------------------------------------------------------------------------------------ localization loss
__global__ void __launch_bounds__(kLossThreads)
    ssdx_localization_loss_kernel(const float4* __restrict__ y_true, const float4* __restrict__ y_pred, const int A,
                                  float* __restrict__ out_loss, float4* __restrict__ out_grad) {
  __shared__ float s_buf[kLossThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float4* yt = y_true + (size_t)b * A;
  const float4* yp = y_pred + (size_t)b * A;
  float sum = 0.0f, cnt = 0.0f;
  for (int a0 = tid; a0 < A; a0 += kLossThreads * 4) {
    float4 t[4], p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int a = a0 + u * kLossThreads;
      const bool ok = a < A;
      t[u] = ok ? (out_grad ? __ldg(yt + a) : ld_stream_f4(yt + a)) : make_float4(0.f, 0.f, 0.f, 0.f);
      p[u] = ok ? (out_grad ? __ldg(yp + a) : ld_stream_f4(yp + a)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float s = __fadd_rn(__fadd_rn(__fadd_rn(fabsf(t[u].x), fabsf(t[u].y)), fabsf(t[u].z)), fabsf(t[u].w));
      if (s > 0.0f) {  // losses.py:23-24 (padding slots are all-zero rows)
        const float d[4] = {__fsub_rn(t[u].x, p[u].x), __fsub_rn(t[u].y, p[u].y), __fsub_rn(t[u].z, p[u].z),
                            __fsub_rn(t[u].w, p[u].w)};
        float row = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float ea = fabsf(d[c]);
          const float l = (ea < 1.0f) ? __fmul_rn(__fmul_rn(d[c], d[c]), 0.5f) : __fsub_rn(ea, 0.5f);  // :27-34
          row = __fadd_rn(row, l);
        }
        sum = __fadd_rn(sum, row);
        cnt = __fadd_rn(cnt, 1.0f);
      }
    }
  }
  const float tot = block_sum(sum, s_buf, tid);
  const float n = block_sum(cnt, s_buf, tid);
  const float den = fmaxf(n, 1.0f);  // :47
  if (tid == 0) out_loss[b] = __fdiv_rn(tot, den);
  if (out_grad == nullptr) return;
  // d loss / d y_pred = -(d inside |d| < 1, sign(d) outside) * not_background / den; second read comes from L2
  float4* g = out_grad + (size_t)b * A;
  for (int a = tid; a < A; a += kLossThreads) {
    const float4 t = __ldg(yt + a), p = __ldg(yp + a);
    const float s = __fadd_rn(__fadd_rn(__fadd_rn(fabsf(t.x), fabsf(t.y)), fabsf(t.z)), fabsf(t.w));
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s > 0.0f) {
      const float d[4] = {__fsub_rn(t.x, p.x), __fsub_rn(t.y, p.y), __fsub_rn(t.z, p.z), __fsub_rn(t.w, p.w)};
      float r[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float sg = (d[c] > 0.0f) ? 1.0f : ((d[c] < 0.0f) ? -1.0f : 0.0f);
        r[c] = __fdiv_rn(-((fabsf(d[c]) < 1.0f) ? d[c] : sg), den);
      }
      o = make_float4(r[0], r[1], r[2], r[3]);
    }
    st_stream_f4(g + a, o);
  }
}

// ------------------------------------------------------------------------------------------ confidence loss
// per-call state in the workspace
struct ConfState {
  unsigned long long n_bg, n_nb;  // count_nonzero(is_background), count_nonzero(not_background) over the batch
  unsigned k_left[3];             // after digit p: elements still to take among the keys that match prefix[p]
  unsigned prefix[3];             // after digit p: decided high bits of the k-th largest key (one slot per digit:
                                  // block 0 of a kernel records slot p while its other blocks still read slot p-1)
  unsigned none;                  // 1 = nothing is kept (no background anchor, or no positive: k = 0)
  unsigned threshold;             // final: the k-th largest key
  unsigned long long ties_take;   // final: how many keys equal to the threshold are kept (lowest flat index first)
  unsigned done;                  // CTAs of the last pass that have finished (the last one writes the losses)
};

struct ConfWs {
  size_t keys, state, hist, l_nb, cnt_nb, sum_gt, tie_cnt, tie_take, total;
};
inline ConfWs conf_ws(int64_t batch, int64_t A) {
  ConfWs w{};
  size_t off = 0;
  w.state = off;    off += ssdx_align_up(sizeof(ConfState), 256);
  w.hist = off;     off += ssdx_align_up((size_t)(kBins0 + kBins1 + kBins2) * sizeof(unsigned), 256);
  w.l_nb = off;     off += ssdx_align_up((size_t)batch * sizeof(float), 256);
  w.cnt_nb = off;   off += ssdx_align_up((size_t)batch * sizeof(float), 256);
  w.sum_gt = off;   off += ssdx_align_up((size_t)batch * sizeof(float), 256);
  w.tie_cnt = off;  off += ssdx_align_up((size_t)batch * sizeof(int), 256);
  w.tie_take = off; off += ssdx_align_up((size_t)batch * sizeof(int), 256);
  w.keys = off;     off += ssdx_align_up((size_t)batch * A * sizeof(unsigned), 256);
  w.total = off;
  return w;
}

// la = -sum_c y_true[c] * log(clip(y_pred[c]))  (losses.py:89-93), any class count
__device__ __forceinline__ float anchor_loss(const float* __restrict__ yt, const float* __restrict__ yp, const int C,
                                             const float lo, const float hi, float& is_bg) {
  float acc = 0.0f;
  is_bg = __ldg(yt);
  for (int c = 0; c < C; ++c) {
    const float t = __ldg(yt + c);
    if (t != 0.0f) acc = __fadd_rn(acc, __fmul_rn(t, logf(fminf(fmaxf(__ldg(yp + c), lo), hi))));  // (zero terms: see below)
  }
  return -acc;
}

// pass 1: per-anchor losses -> keys of the background part + first histogram; per-image sums of the positive part
template <int CV>
__global__ void __launch_bounds__(kLossThreads)
    ssdx_confidence_pass1_kernel(const float* __restrict__ y_true, const float* __restrict__ y_pred, const int A,
                                 const int C, const float lo, const float hi, unsigned* __restrict__ keys,
                                 unsigned* __restrict__ hist0, float* __restrict__ l_nb, float* __restrict__ cnt_nb,
                                 ConfState* __restrict__ st) {
  __shared__ unsigned s_hist[kBins0];
  __shared__ float s_buf[kLossThreads / 32];
  __shared__ int s_cnt[2];
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < kBins0; i += kLossThreads) s_hist[i] = 0u;
  if (tid < 2) s_cnt[tid] = 0;
  __syncthreads();
  const float* yt = y_true + (size_t)b * A * C;
  const float* yp = y_pred + (size_t)b * A * C;
  unsigned* kb = keys + (size_t)b * A;
  float sum_nb = 0.0f, n_nb_f = 0.0f;
  int c_bg = 0, c_nb = 0;
  constexpr int U = (CV == 4) ? 4 : 1;  // anchors in flight per thread (2 U independent 128-bit loads)
  for (int base = 0; base < A; base += kLossThreads * U) {
    const int a0 = base + tid;
    float la[U], is_bg[U];
    if (CV == 4) {
      float4 t[U], p[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int a = a0 + u * kLossThreads;
        const bool ok = a < A;
        t[u] = ok ? ld_stream_f4(reinterpret_cast<const float4*>(yt) + a) : make_float4(0.f, 0.f, 0.f, 0.f);
        p[u] = ok ? ld_stream_f4(reinterpret_cast<const float4*>(yp) + a) : make_float4(1.f, 1.f, 1.f, 1.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        // a term with a zero target is +-0 (the clipped logarithm is finite) and leaves the running sum unchanged
        // bit for bit, so its logarithm is skipped: one logf per one-hot row instead of four (the four full-precision
        // logarithms made this pass instruction-bound: 73 us against 58 us)
        float acc = 0.0f;
        if (t[u].x != 0.0f) acc = __fadd_rn(acc, __fmul_rn(t[u].x, logf(fminf(fmaxf(p[u].x, lo), hi))));
        if (t[u].y != 0.0f) acc = __fadd_rn(acc, __fmul_rn(t[u].y, logf(fminf(fmaxf(p[u].y, lo), hi))));
        if (t[u].z != 0.0f) acc = __fadd_rn(acc, __fmul_rn(t[u].z, logf(fminf(fmaxf(p[u].z, lo), hi))));
        if (t[u].w != 0.0f) acc = __fadd_rn(acc, __fmul_rn(t[u].w, logf(fminf(fmaxf(p[u].w, lo), hi))));
        la[u] = -acc;
        is_bg[u] = t[u].x;
      }
    } else {
      la[0] = is_bg[0] = 0.0f;
      if (a0 < A) la[0] = anchor_loss(yt + (size_t)a0 * C, yp + (size_t)a0 * C, C, lo, hi, is_bg[0]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int a = a0 + u * kLossThreads;
      const bool ok = a < A;
      unsigned bin = 0xFFFFFFFFu;
      if (ok) {
        const float not_bg = fabsf(__fsub_rn(is_bg[u], 1.0f));           // :70
        sum_nb = __fadd_rn(sum_nb, __fmul_rn(la[u], not_bg));            // :100-101
        n_nb_f = __fadd_rn(n_nb_f, not_bg);                              // :104
        c_bg += (is_bg[u] != 0.0f);                                      // :73
        c_nb += (not_bg != 0.0f);                                        // :74
        const unsigned key = loss_key(__fmul_rn(la[u], is_bg[u]));       // :120
        kb[a] = key;
        bin = key >> 21;
      }
      if (ok) atomicAdd(&s_hist[bin], 1u);  // (warp-aggregating with match_any first measured slower: 91 vs 66 us)
    }
  }
  const float tot = block_sum(sum_nb, s_buf, tid);
  const float n = block_sum(n_nb_f, s_buf, tid);
  c_bg = __reduce_add_sync(0xFFFFFFFFu, c_bg);
  c_nb = __reduce_add_sync(0xFFFFFFFFu, c_nb);
  if ((tid & 31) == 0) {
    atomicAdd(&s_cnt[0], c_bg);
    atomicAdd(&s_cnt[1], c_nb);
  }
  __syncthreads();
  if (tid == 0) {
    l_nb[b] = tot;
    cnt_nb[b] = n;
    atomicAdd(&st->n_bg, (unsigned long long)s_cnt[0]);
    atomicAdd(&st->n_nb, (unsigned long long)s_cnt[1]);
  }
  for (int i = tid; i < kBins0; i += kLossThreads) {
    const unsigned v = s_hist[i];
    if (v) atomicAdd(hist0 + i, v);
  }
}

// Every block walks the histogram of the previous digit from the top and finds the bin that holds the k-th largest key
// (redundantly: 2048 counters from L2 and one block scan cost less than a launch); block 0 records the result.
struct Pick {
  unsigned prefix, k_left, none;
};
template <int BINS, int SHIFT, int PASS, int THREADS>
__device__ __forceinline__ Pick pick_digit(const unsigned* __restrict__ hist, ConfState* __restrict__ st,
                                           const bool record) {
  using Scan = cub::BlockScan<unsigned, THREADS>;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ Pick s_pick;
  const int tid = threadIdx.x;
  unsigned k, prefix_in = 0u;
  if (PASS == 0) {
    const unsigned long long n_bg = st->n_bg, n_nb = st->n_nb;
    k = (unsigned)((3ull * n_nb < n_bg) ? 3ull * n_nb : n_bg);  // :117
    if (n_bg == 0ull || k == 0u) {                              // :123 / top_k with k = 0: nothing is kept
      if (record && tid == 0) {
        st->none = 1u;
        st->threshold = 0xFFFFFFFFu;
        st->ties_take = 0ull;
      }
      return Pick{0u, 0u, 1u};
    }
  } else {
    if (st->none) return Pick{0u, 0u, 1u};
    k = st->k_left[PASS - 1];
    prefix_in = st->prefix[PASS - 1];
  }
  constexpr int PER = BINS / THREADS;
  static_assert(BINS % THREADS == 0, "bins per thread");
  unsigned c[PER], tot = 0u;  // thread t owns bins BINS-1 - (t*PER + j), j = 0..PER-1 (descending)
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    c[j] = __ldg(hist + (BINS - 1 - (tid * PER + j)));
    tot += c[j];
  }
  unsigned before;
  Scan(tmp).ExclusiveSum(tot, before);
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    if (before < k && k <= before + c[j]) {  // exactly one (thread, j)
      s_pick.prefix = prefix_in | ((unsigned)(BINS - 1 - (tid * PER + j)) << SHIFT);
      s_pick.k_left = k - before;  // rank inside the bin
      s_pick.none = 0u;
    }
    before += c[j];
  }
  __syncthreads();
  const Pick p = s_pick;
  if (record && tid == 0) {
    st->prefix[PASS] = p.prefix;
    st->k_left[PASS] = p.k_left;
    if (PASS == 2) {
      st->threshold = p.prefix;
      st->ties_take = p.k_left;
    }
  }
  return p;
}

// histogram of the next digit over the keys that match the prefix decided so far (PASS = index of the digit that
// this kernel picks first: 0 -> histogram of digit 1, 1 -> histogram of digit 2)
template <int PASS>
__global__ void __launch_bounds__(kHistThreads) ssdx_confidence_hist_kernel(const unsigned* __restrict__ keys,
                                                                            const long long n, unsigned* __restrict__ hist_all,
                                                                            ConfState* __restrict__ st) {
  constexpr int BINS = PASS == 0 ? kBins1 : kBins2;
  constexpr int SHIFT = PASS == 0 ? 10 : 0;
  constexpr unsigned PREFIX_MASK = PASS == 0 ? 0xFFE00000u : 0xFFFFFC00u;
  __shared__ unsigned s_hist[BINS];
  const Pick pk = (PASS == 0) ? pick_digit<kBins0, 21, 0, kHistThreads>(hist_all, st, blockIdx.x == 0)
                              : pick_digit<kBins1, 10, 1, kHistThreads>(hist_all + kBins0, st, blockIdx.x == 0);
  if (pk.none) return;
  unsigned* hist = hist_all + (PASS == 0 ? kBins0 : kBins0 + kBins1);
  const unsigned prefix = pk.prefix;
  for (int i = threadIdx.x; i < BINS; i += kHistThreads) s_hist[i] = 0u;
  __syncthreads();
  const long long n4 = n >> 2;
  const uint4* k4 = reinterpret_cast<const uint4*>(keys);
  const long long stride = (long long)gridDim.x * kHistThreads;
  for (long long i0 = (long long)blockIdx.x * kHistThreads + threadIdx.x; i0 < n4; i0 += 2 * stride) {
    const long long i1 = i0 + stride;
    const uint4 v = __ldg(k4 + i0);
    const uint4 w = (i1 < n4) ? __ldg(k4 + i1) : make_uint4(~prefix, ~prefix, ~prefix, ~prefix);
    if ((v.x & PREFIX_MASK) == prefix) atomicAdd(&s_hist[(v.x >> SHIFT) & (BINS - 1)], 1u);
    if ((v.y & PREFIX_MASK) == prefix) atomicAdd(&s_hist[(v.y >> SHIFT) & (BINS - 1)], 1u);
    if ((v.z & PREFIX_MASK) == prefix) atomicAdd(&s_hist[(v.z >> SHIFT) & (BINS - 1)], 1u);
    if ((v.w & PREFIX_MASK) == prefix) atomicAdd(&s_hist[(v.w >> SHIFT) & (BINS - 1)], 1u);
    if ((w.x & PREFIX_MASK) == prefix) atomicAdd(&s_hist[(w.x >> SHIFT) & (BINS - 1)], 1u);
    if ((w.y & PREFIX_MASK) == prefix) atomicAdd(&s_hist[(w.y >> SHIFT) & (BINS - 1)], 1u);
    if ((w.z & PREFIX_MASK) == prefix) atomicAdd(&s_hist[(w.z >> SHIFT) & (BINS - 1)], 1u);
    if ((w.w & PREFIX_MASK) == prefix) atomicAdd(&s_hist[(w.w >> SHIFT) & (BINS - 1)], 1u);
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += kHistThreads) {
      const unsigned v = __ldg(keys + i);
      if ((v & PREFIX_MASK) == prefix) atomicAdd(&s_hist[(v >> SHIFT) & (BINS - 1)], 1u);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < BINS; i += kHistThreads) {
    const unsigned v = s_hist[i];
    if (v) atomicAdd(hist + i, v);
  }
}

// per image: sum of the background losses above the threshold, number of keys equal to it.  The CTA that finishes
// last turns the per-image tie counts into shares (flat index order) and writes the loss values.
__global__ void __launch_bounds__(kLossThreads)
    ssdx_confidence_above_kernel(const unsigned* __restrict__ keys, const int A, const int batch,
                                 const unsigned* __restrict__ hist_all, ConfState* __restrict__ st,
                                 const float* __restrict__ l_nb, const float* __restrict__ cnt_nb,
                                 float* __restrict__ sum_gt, int* __restrict__ tie_cnt, int* __restrict__ tie_take,
                                 unsigned* __restrict__ done, float* __restrict__ out_loss) {
  __shared__ float s_buf[kLossThreads / 32];
  __shared__ int s_t;
  __shared__ bool s_last;
  const int b = blockIdx.x, tid = threadIdx.x;
  const Pick pk = pick_digit<kBins2, 0, 2, kLossThreads>(hist_all + kBins0 + kBins1, st, b == 0);
  const bool none = pk.none != 0u;
  const unsigned T = none ? 0xFFFFFFFFu : pk.prefix;
  if (tid == 0) s_t = 0;
  const unsigned* kb = keys + (size_t)b * A;
  float sum = 0.0f;
  int ties = 0;
  auto visit = [&](const unsigned k) {
    if (k > T) sum = __fadd_rn(sum, key_loss(k));
    ties += (k == T);
  };
  if (!none) {
    // 16-byte loads over the aligned middle of the row, scalar head and tail
    const int head = min(A, (int)((4 - (((size_t)b * A) & 3)) & 3));
    const int n4 = (A - head) >> 2;
    const uint4* k4 = reinterpret_cast<const uint4*>(kb + head);
    for (int i = tid; i < n4; i += kLossThreads) {
      const uint4 v = __ldg(k4 + i);
      visit(v.x); visit(v.y); visit(v.z); visit(v.w);
    }
    if (tid < head) visit(__ldg(kb + tid));
    for (int a = head + (n4 << 2) + tid; a < A; a += kLossThreads) visit(__ldg(kb + a));
  }
  const float tot = block_sum(sum, s_buf, tid);  // (its barriers also order s_t = 0)
  ties = __reduce_add_sync(0xFFFFFFFFu, ties);
  if ((tid & 31) == 0 && ties) atomicAdd(&s_t, ties);
  __syncthreads();
  if (tid == 0) {
    sum_gt[b] = tot;
    tie_cnt[b] = s_t;
    __threadfence();
    s_last = atomicAdd(done, 1u) == (unsigned)gridDim.x - 1u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- last CTA: exclusive scan of the tie counts over the images -> share of each image, loss values ----
  using Scan = cub::BlockScan<unsigned, kLossThreads>;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ unsigned s_carry;
  if (tid == 0) s_carry = 0u;
  __syncthreads();
  const unsigned take_all = none ? 0u : pk.k_left;
  const float tval = key_loss(T);
  const volatile int* v_cnt = tie_cnt;
  const volatile float* v_sum = sum_gt;
  for (int b0 = 0; b0 < batch; b0 += kLossThreads) {
    const int i = b0 + tid;
    const unsigned mine = (i < batch && !none) ? (unsigned)v_cnt[i] : 0u;
    unsigned before, chunk;
    Scan(tmp).ExclusiveSum(mine, before, chunk);
    before += s_carry;
    __syncthreads();
    if (tid == 0) s_carry += chunk;
    if (i < batch) {
      const unsigned left = take_all > before ? take_all - before : 0u;
      const int take = (int)(left < mine ? left : mine);
      tie_take[i] = take;
      float l_bg = 0.0f;
      if (!none) l_bg = __fadd_rn(v_sum[i], __fmul_rn((float)take, tval));          // :156-159
      out_loss[i] = __fdiv_rn(__fadd_rn(l_nb[i], l_bg), fmaxf(cnt_nb[i], 1.0f));   // :164-172
    }
    __syncthreads();
  }
}

// gradient and / or the kept-background mask: one CTA per image, anchors visited in index order so that the image's
// share of the threshold ties goes to its lowest anchors
template <int CV>
__global__ void __launch_bounds__(kLossThreads)
    ssdx_confidence_grad_kernel(const float* __restrict__ y_true, const float* __restrict__ y_pred, const int A,
                                const int C, const float lo, const float hi, const unsigned* __restrict__ keys,
                                const ConfState* __restrict__ st, const float* __restrict__ cnt_nb,
                                const int* __restrict__ tie_cnt, const int* __restrict__ tie_take,
                                float* __restrict__ out_grad, unsigned char* __restrict__ out_keep) {
  __shared__ int s_warp[kLossThreads / 32];
  __shared__ int s_base;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool none = st->none != 0u;
  const unsigned T = st->threshold;
  const int take = none ? 0 : tie_take[b];
  const float den = fmaxf(cnt_nb[b], 1.0f);
  const float* yt = y_true + (size_t)b * A * C;
  const float* yp = y_pred + (size_t)b * A * C;
  const unsigned* kb = keys + (size_t)b * A;
  // keep-or-drop of one anchor -> gradient row and mask byte
  auto emit = [&](const int a, const bool keep) {
    if (out_keep) out_keep[(size_t)b * A + a] = keep ? 1 : 0;
    if (out_grad == nullptr) return;
    const float* t = yt + (size_t)a * C;
    const float* p = yp + (size_t)a * C;
    float* g = out_grad + ((size_t)b * A + a) * C;
    if (CV == 4) {
      const float4 tv = ld_stream_f4(reinterpret_cast<const float4*>(t));
      const float4 pv = ld_stream_f4(reinterpret_cast<const float4*>(p));
      const float is_bg = tv.x, not_bg = fabsf(__fsub_rn(tv.x, 1.0f));
      const float w = -__fdiv_rn(__fadd_rn(not_bg, keep ? is_bg : 0.0f), den);
      const float tt[4] = {tv.x, tv.y, tv.z, tv.w}, pp[4] = {pv.x, pv.y, pv.z, pv.w};
      float r[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const bool pass = pp[c] >= lo && pp[c] <= hi && tt[c] != 0.0f;  // zero target: the quotient is +-0, skipped
        r[c] = pass ? __fdiv_rn(__fmul_rn(w, tt[c]), fminf(fmaxf(pp[c], lo), hi)) : 0.0f;
      }
      st_stream_f4(reinterpret_cast<float4*>(g), make_float4(r[0], r[1], r[2], r[3]));
    } else {
      const float is_bg = __ldg(t), not_bg = fabsf(__fsub_rn(is_bg, 1.0f));
      const float w = -__fdiv_rn(__fadd_rn(not_bg, keep ? is_bg : 0.0f), den);
      for (int c = 0; c < C; ++c) {
        const float pc = __ldg(p + c), tc = __ldg(t + c);
        const bool pass = pc >= lo && pc <= hi && tc != 0.0f;
        g[c] = pass ? __fdiv_rn(__fmul_rn(w, tc), fminf(fmaxf(pc, lo), hi)) : 0.0f;
      }
    }
  };
  const int n_ties = none ? 0 : tie_cnt[b];
  if (take == 0 || take == n_ties) {
    // every image but (at most) one: its threshold ties are all kept or all dropped, no ranking needed
    const bool ties_kept = take > 0;
#pragma unroll 4
    for (int a = tid; a < A; a += kLossThreads) {
      const unsigned key = __ldg(kb + a);
      emit(a, !none && (key > T || (ties_kept && key == T)));
    }
    return;
  }
  // the image where the share of ties runs out: rank the ties in index order (ballot inside the warp, prefix over
  // the warps, running base over the chunks)
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int a0 = 0; a0 < A; a0 += kLossThreads) {
    const int a = a0 + tid;
    const unsigned key = (a < A) ? __ldg(kb + a) : 0u;
    const bool tie = a < A && key == T;
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, tie);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int rank = s_base + __popc(bal & ((1u << lane) - 1u));
    int chunk = 0;
    for (int w = 0; w < kLossThreads / 32; ++w) {
      const int c = s_warp[w];
      if (w < warp) rank += c;
      chunk += c;
    }
    __syncthreads();
    if (tid == 0) s_base += chunk;
    if (a < A) emit(a, key > T || (tie && rank < take));
    __syncthreads();
  }
}

}  // namespace

extern "C" int ssdx_localization_loss(const float* y_true, const float* y_pred, int32_t batch, int32_t num_anchors,
                                      float* out_loss, float* out_grad, int device, ssdx_stream stream) {
  SSDX_CHECK_ARG(y_true != nullptr && y_pred != nullptr && out_loss != nullptr,
                 "ssdx_localization_loss: NULL pointer argument");
  SSDX_CHECK_ARG(batch >= 0 && num_anchors > 0, "ssdx_localization_loss: batch must be >= 0 and num_anchors > 0");
  SSDX_CHECK_ARG(((uintptr_t)y_true & 15) == 0 && ((uintptr_t)y_pred & 15) == 0 && ((uintptr_t)out_grad & 15) == 0,
                 "ssdx_localization_loss: y_true / y_pred / out_grad must be 16-byte aligned");
  if (batch == 0) return SSDX_OK;
  SSDX_DEVICE(device);
  ssdx_localization_loss_kernel<<<batch, kLossThreads, 0, (cudaStream_t)stream>>>(
      (const float4*)y_true, (const float4*)y_pred, num_anchors, out_loss, (float4*)out_grad);
  SSDX_LAUNCH_CHECK("ssdx_localization_loss_kernel");
  return SSDX_OK;
}

extern "C" size_t ssdx_confidence_loss_workspace_bytes(int32_t batch, int32_t num_anchors) {
  if (batch < 0 || num_anchors < 0) return 0;
  return conf_ws(batch, num_anchors).total;
}

extern "C" int ssdx_confidence_loss(const float* y_true, const float* y_pred, int32_t batch, int32_t num_anchors,
                                    int32_t num_classes, float clip_min, float clip_max, float* out_loss,
                                    float* out_grad, uint8_t* out_keep, void* workspace, size_t workspace_bytes,
                                    int device, ssdx_stream stream) {
  SSDX_CHECK_ARG(y_true != nullptr && y_pred != nullptr && out_loss != nullptr,
                 "ssdx_confidence_loss: NULL pointer argument");
  SSDX_CHECK_ARG(batch >= 0 && num_anchors > 0 && num_classes > 0,
                 "ssdx_confidence_loss: batch must be >= 0, num_anchors and num_classes > 0");
  SSDX_CHECK_ARG((long long)batch * num_anchors < (1ll << 31), "ssdx_confidence_loss: batch * num_anchors must be < 2^31");
  if (batch == 0) return SSDX_OK;
  const ConfWs w = conf_ws(batch, num_anchors);
  if (workspace == nullptr || workspace_bytes < w.total) {
    ssdx_set_error("ssdx_confidence_loss: workspace too small (%zu < %zu bytes)", workspace_bytes, w.total);
    return SSDX_ERR_WORKSPACE;
  }
  SSDX_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "ssdx_confidence_loss: workspace must be 256-byte aligned");
  const bool vec4 = num_classes == 4 && ((uintptr_t)y_true & 15) == 0 && ((uintptr_t)y_pred & 15) == 0 &&
                    ((uintptr_t)out_grad & 15) == 0;
  SSDX_DEVICE(device);
  cudaStream_t s = (cudaStream_t)stream;
  char* base = (char*)workspace;
  ConfState* st = (ConfState*)(base + w.state);
  unsigned* hist = (unsigned*)(base + w.hist);
  unsigned* keys = (unsigned*)(base + w.keys);
  float* l_nb = (float*)(base + w.l_nb);
  float* cnt_nb = (float*)(base + w.cnt_nb);
  float* sum_gt = (float*)(base + w.sum_gt);
  int* tie_cnt = (int*)(base + w.tie_cnt);
  int* tie_take = (int*)(base + w.tie_take);
  const int A = num_anchors, C = num_classes;
  const long long n = (long long)batch * A;
  SSDX_CUDA(cudaMemsetAsync(base, 0, w.l_nb, s));  // state + the three histograms
  if (vec4)
    ssdx_confidence_pass1_kernel<4><<<batch, kLossThreads, 0, s>>>(y_true, y_pred, A, C, clip_min, clip_max, keys, hist,
                                                                   l_nb, cnt_nb, st);
  else
    ssdx_confidence_pass1_kernel<0><<<batch, kLossThreads, 0, s>>>(y_true, y_pred, A, C, clip_min, clip_max, keys, hist,
                                                                   l_nb, cnt_nb, st);
  SSDX_LAUNCH_CHECK("ssdx_confidence_pass1_kernel");
  const int hist_grid = 4 * ssdx_sm_count(device);
  ssdx_confidence_hist_kernel<0><<<hist_grid, kHistThreads, 0, s>>>(keys, n, hist, st);
  ssdx_confidence_hist_kernel<1><<<hist_grid, kHistThreads, 0, s>>>(keys, n, hist, st);
  SSDX_LAUNCH_CHECK("ssdx_confidence_hist_kernel");
  ssdx_confidence_above_kernel<<<batch, kLossThreads, 0, s>>>(keys, A, batch, hist, st, l_nb, cnt_nb, sum_gt, tie_cnt,
                                                              tie_take, &st->done, out_loss);
  SSDX_LAUNCH_CHECK("ssdx_confidence_above_kernel");
  if (out_grad != nullptr || out_keep != nullptr) {
    if (vec4)
      ssdx_confidence_grad_kernel<4><<<batch, kLossThreads, 0, s>>>(y_true, y_pred, A, C, clip_min, clip_max, keys, st,
                                                                    cnt_nb, tie_cnt, tie_take, out_grad, out_keep);
    else
      ssdx_confidence_grad_kernel<0><<<batch, kLossThreads, 0, s>>>(y_true, y_pred, A, C, clip_min, clip_max, keys, st,
                                                                    cnt_nb, tie_cnt, tie_take, out_grad, out_keep);
    SSDX_LAUNCH_CHECK("ssdx_confidence_grad_kernel");
  }
  return SSDX_OK;
}
