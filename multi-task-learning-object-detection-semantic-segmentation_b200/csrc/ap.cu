// Average precision on the device (reference: evaluators.py:159-187), SURVEY 8(f) row 4: the tail of
// average_precision_object_detection after the TP flags.  Per foreground class: predictions sorted by confidence
// (descending), cum = cumsum(tp) (fp32, exact: counts stay below 2^24), precision = cum / rank (fp64),
// recall = cum / n_gt (fp32, as NumPy computes float32 / int), AP = trapezoid(precision over recall) in fp64 with no
// (0, .) sentinel; 0 when the class has no ground truth or no prediction.
//
// Recall only moves on a true positive, so only the TP predictions contribute a trapezoid, and the term of a TP
// prediction needs nothing but its rank r in its class and the number cum of TPs up to it:
//     term = (cum / n_gt - (cum - 1) / n_gt) * ((cum - 1) / r + cum / (r + 1)) / 2        (r >= 1).
// Ranks are counted, nothing is sorted into place (a library radix sort spent 8 latency-bound passes on the few
// hundred thousand predictions of an evaluation: 470 us for 819,200 of them):
//   every prediction has the 64-bit composite  class << 56 | ~confidence key << 24 | index  (all distinct: ties in
//   confidence fall in input order - NumPy's quicksort leaves them unspecified - and any sample splits evenly);
//   A  one CTA sorts a strided sample of 4096 composites (4 keys per thread: registers, shuffles, 25 shared-memory
//      exchanges); every (4096 / S)-th one is a splitter;
//   B  bucket of a prediction = (number of splitters <= composite) + class: monotone in the composite and one class
//      per bucket; per (CTA, bucket) counts and per-bucket TP counts - every CTA owns a contiguous chunk;
//   C  column scans over the CTAs (one warp per bucket), then one CTA: bucket offsets and per bucket the rank / TP
//      base inside its class;
//   D  the chunks are scattered into the buckets (cursors in shared memory; order inside a bucket is irrelevant);
//      the scattered key drops the class and carries the TP flag:  (~confidence key << 24 | index) << 1 | tp;
//   E  one CTA per bucket (~1000 keys in shared memory): keys are spread over 256 equal slices of the bucket's key
//      range (a counting pass), and every TP key counts the keys / TP keys below it inside its slice - the rank and
//      the TP count follow from the slice and bucket bases; fp64 terms, one partial per bucket;
//   F  per class, the partials in a fixed order.
// A bucket that outgrows shared memory, or a slice that holds many keys (thousands of equal confidences), only costs
// time: the counts are exact either way.  More than 2^24 - 1 predictions or 254 classes do not fit the composite and
// are refused.
#include "common.cuh"

namespace {

typedef unsigned long long u64;
constexpr int kApThreads = 512;
constexpr int kSample = 4096;        // composites sampled for the splitters
constexpr int kBucketTarget = 1024;  // predictions per bucket aimed at
constexpr int kBucketCap = 6144;     // keys of a bucket held in shared memory (twice: as scattered, and by slice): 6x the target
constexpr int kMaxSplit = 1024;      // sample splitters at most
constexpr int kMaxBuckets = kMaxSplit + 256;
constexpr int kSlices = 256;
constexpr int kApCtas = 148;         // CTAs (= contiguous chunks of the predictions) of the bucket / scatter passes
static_assert(kApCtas <= 160, "the column scan of ssdx_ap_offsets_kernel holds 5 CTAs per lane");
constexpr unsigned kUnscored = 0xFFu;
constexpr u64 kPad = ~0ull;

__device__ __forceinline__ unsigned conf_key_desc(float s) {  // smaller key = larger confidence
  const unsigned u = __float_as_uint(s);
  const unsigned asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~asc;
}
__device__ __forceinline__ unsigned scored_class(const int l, const int C, const int bg) {
  return (l != bg && l >= 0 && l < C) ? (unsigned)l : kUnscored;
}
__device__ __forceinline__ u64 composite(const unsigned cls, const float conf, const long long i) {
  return ((u64)cls << 56) | ((u64)conf_key_desc(conf) << 24) | (u64)i;
}

// bucket = number of splitters <= key
__device__ __forceinline__ int splitters_below(const u64* splitters, const int n_split, const u64 key) {
  int lo = 0, hi = n_split;  // first index with splitters[idx] > key
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (splitters[mid] <= key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// A: splitters[k] = sorted sample[(k + 1) * ns / S], k < S - 1.  The 4096 sampled composites are sorted (ascending)
// as 4 keys per thread of one 1024-thread CTA: exchange distances of 1024 and 2048 stay inside a thread, distances
// below 32 are warp shuffles, only the 25 stages at distances 32..512 go through shared memory (a smem-only network
// takes a barrier in each of its 78 stages: 41 us, the longest kernel of the AP tail).
__global__ void __launch_bounds__(1024)
    ssdx_ap_sample_kernel(const float* __restrict__ conf, const int* __restrict__ label, const long long n, const int C,
                          const int bg, const int S, u64* __restrict__ splitters) {
  static_assert(kSample == 4 * 1024, "4 keys per thread");
  __shared__ u64 s[kSample];
  const int tid = threadIdx.x;
  const int ns = (int)(n < kSample ? n : kSample);
  u64 v[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {  // key e = r * 1024 + tid
    const int j = r * 1024 + tid;
    v[r] = kPad;
    if (j < ns) {
      const long long i = (long long)j * n / ns;
      v[r] = composite(scored_class(label[i], C, bg), conf[i], i);
    }
  }
#pragma unroll
  for (int k = 2; k <= kSample; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 1024) {  // both keys of a pair live in this thread
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int rp = r ^ (j >> 10);
          if (rp > r) {
            const bool up = (((r * 1024 + tid) & k) == 0);  // ascending block: the smaller key at the lower index
            const u64 a = v[r], b = v[rp];
            const bool swap = up ? (a > b) : (a < b);
            v[r] = swap ? b : a;
            v[rp] = swap ? a : b;
          }
        }
      } else if (j < 32) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int e = r * 1024 + tid;
          const u64 p = __shfl_xor_sync(0xFFFFFFFFu, v[r], j);
          const bool take_min = (((e & j) == 0) == ((e & k) == 0));
          v[r] = ((v[r] < p) == take_min) ? v[r] : p;  // (equal keys: either copy)
        }
      } else {
        __syncthreads();  // the previous exchange has been read
#pragma unroll
        for (int r = 0; r < 4; ++r) s[r * 1024 + tid] = v[r];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int e = r * 1024 + tid;
          const u64 p = s[e ^ j];
          const bool take_min = (((e & j) == 0) == ((e & k) == 0));
          v[r] = ((v[r] < p) == take_min) ? v[r] : p;
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) s[r * 1024 + tid] = v[r];
  __syncthreads();
  for (int k = tid; k < S - 1; k += 1024) splitters[k] = s[(int)(((long long)(k + 1) * ns) / S)];
}

// B: bucket ids; per (CTA, bucket) counts and per-bucket TP counts (tpc: zeroed by the host; bcls: preset to -1)
__global__ void __launch_bounds__(1024)
    ssdx_ap_bucket_kernel(const unsigned char* __restrict__ tp, const float* __restrict__ conf,
                          const int* __restrict__ label, const long long n, const int C, const int bg,
                          const u64* __restrict__ splitters, const int n_split, const int nb,
                          unsigned short* __restrict__ bid, int* __restrict__ cta_cnt, int* __restrict__ tpc,
                          int* __restrict__ bcls) {
  __shared__ u64 s_sp[kMaxSplit];
  __shared__ int s_cnt[kMaxBuckets], s_tp[kMaxBuckets];
  for (int k = threadIdx.x; k < n_split; k += blockDim.x) s_sp[k] = splitters[k];
  for (int k = threadIdx.x; k < nb; k += blockDim.x) {
    s_cnt[k] = 0;
    s_tp[k] = 0;
  }
  __syncthreads();
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long lo = per * blockIdx.x, hi = (lo + per < n) ? lo + per : n;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const unsigned cls = scored_class(label[i], C, bg);
    unsigned short b = 0xFFFFu;
    if (cls != kUnscored) {
      b = (unsigned short)(splitters_below(s_sp, n_split, composite(cls, conf[i], i)) + (int)cls);
      if (atomicAdd(&s_cnt[b], 1) == 0) bcls[b] = (int)cls;  // (every writer of a bucket writes the same class)
      if (tp[i]) atomicAdd(&s_tp[b], 1);
    }
    bid[i] = b;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nb; k += blockDim.x) {
    cta_cnt[(size_t)k * kApCtas + blockIdx.x] = s_cnt[k];  // [bucket][CTA]: the column scan reads rows of this
    if (s_tp[k]) atomicAdd(tpc + k, s_tp[k]);
  }
}

// C1: per bucket (one warp each), the exclusive prefix of its counts over the CTAs (kApCtas <= 160 = 5 per lane) and
// the bucket total
__global__ void __launch_bounds__(256)
    ssdx_ap_columns_kernel(int* __restrict__ cta_cnt, const int n_ctas, const int nb, int* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= nb) return;
  int* row = cta_cnt + (size_t)b * kApCtas;
  int v[5], sum = 0;
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const int k = lane * 5 + q;
    v[q] = k < n_ctas ? row[k] : 0;
    sum += v[q];
  }
  int x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
    if (lane >= o) x += y;
  }
  int run = x - sum;
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const int k = lane * 5 + q;
    if (k < n_ctas) row[k] = run;
    run += v[q];
  }
  if (lane == 31) cnt[b] = x;
}

// C2: one CTA.  Bucket offsets, and per bucket the rank / cumulative-TP base inside its class (buckets are
// class-sorted: base = running total minus the running total at the first bucket of the class).
__global__ void __launch_bounds__(1024)
    ssdx_ap_offsets_kernel(const int nb, const int* __restrict__ tpc, const int* __restrict__ bcls,
                           const int* __restrict__ cnt, int* __restrict__ off, int* __restrict__ rank_base,
                           int* __restrict__ tp_base) {
  __shared__ int s_off[kMaxBuckets + 1], s_tpo[kMaxBuckets + 1];
  __shared__ int s_first_off[256], s_first_tp[256];
  __shared__ int s_wa[32], s_wb[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 256) {
    s_first_off[tid] = 0x7FFFFFFF;
    s_first_tp[tid] = 0x7FFFFFFF;
  }
  // two buckets per thread: b = 2 * tid, 2 * tid + 1 (kMaxBuckets <= 2048)
  int c2[2], t2[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int b = 2 * tid + u;
    c2[u] = b < nb ? cnt[b] : 0;
    t2[u] = b < nb ? tpc[b] : 0;
  }
  int xa = c2[0] + c2[1], xb = t2[0] + t2[1];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int ya = __shfl_up_sync(0xFFFFFFFFu, xa, o), yb = __shfl_up_sync(0xFFFFFFFFu, xb, o);
    if (lane >= o) {
      xa += ya;
      xb += yb;
    }
  }
  if (lane == 31) {
    s_wa[warp] = xa;
    s_wb[warp] = xb;
  }
  __syncthreads();
  int ba = 0, bb = 0;
  for (int w = 0; w < warp; ++w) {
    ba += s_wa[w];
    bb += s_wb[w];
  }
  int oa = ba + xa - (c2[0] + c2[1]), ob = bb + xb - (t2[0] + t2[1]);  // exclusive prefixes at bucket 2 * tid
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int b = 2 * tid + u;
    if (b < nb) {
      s_off[b] = oa;
      s_tpo[b] = ob;
      const int cls = bcls[b];
      if (cls >= 0 && cls < 256 && c2[u] > 0) {
        atomicMin(&s_first_off[cls], oa);  // prefixes are non-decreasing: the minimum is the class's first bucket
        atomicMin(&s_first_tp[cls], ob);
      }
    }
    oa += c2[u];
    ob += t2[u];
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int b = 2 * tid + u;
    if (b < nb) {
      const int cls = bcls[b];
      const bool live = cls >= 0 && cls < 256 && c2[u] > 0;
      off[b] = s_off[b];
      rank_base[b] = live ? s_off[b] - s_first_off[cls] : 0;
      tp_base[b] = live ? s_tpo[b] - s_first_tp[cls] : 0;
    }
  }
}

// D: scatter.  CTA k places its chunk: bucket b starts at off[b] + (prefix of the CTAs before k); cursors in shared
// memory.  Scattered key = (~confidence key << 24 | index) << 1 | tp   (the class is the bucket's).
__global__ void __launch_bounds__(512)
    ssdx_ap_scatter_kernel(const unsigned char* __restrict__ tp, const float* __restrict__ conf,
                           const unsigned short* __restrict__ bid, const long long n, const int nb,
                           const int* __restrict__ off, const int* __restrict__ cta_cnt, u64* __restrict__ keys) {
  __shared__ int s_cur[kMaxBuckets];
  for (int k = threadIdx.x; k < nb; k += blockDim.x) s_cur[k] = off[k] + cta_cnt[(size_t)k * kApCtas + blockIdx.x];
  __syncthreads();
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long lo = per * blockIdx.x, hi = (lo + per < n) ? lo + per : n;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const unsigned short b = bid[i];
    if (b == 0xFFFFu) continue;
    keys[atomicAdd(&s_cur[b], 1)] = ((((u64)conf_key_desc(conf[i]) << 24) | (u64)i) << 1) | (u64)(tp[i] != 0);
  }
}

// E: one CTA per bucket
__global__ void __launch_bounds__(kApThreads)
    ssdx_ap_terms_kernel(const u64* __restrict__ keys, const int* __restrict__ cnt, const int* __restrict__ off,
                         const int* __restrict__ rank_base, const int* __restrict__ tp_base,
                         const int* __restrict__ bcls, const long long* __restrict__ counts /*C,3*/, const int C,
                         const int bg, double* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char ap_smem[];
  u64* s_in = reinterpret_cast<u64*>(ap_smem);  // [kBucketCap] as scattered
  u64* s_by = s_in + kBucketCap;                // [kBucketCap] grouped by slice
  __shared__ int s_scnt[kSlices + 1], s_stp[kSlices + 1], s_cur[kSlices];
  __shared__ u64 s_min, s_max;
  __shared__ double s_red[kApThreads / 32];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = cnt[b];
  const int cls = bcls[b];
  double acc = 0.0;
  const long long n_gt = (cls >= 0 && cls < C) ? counts[(size_t)cls * 3 + 2] : 0;
  if (m > 0 && cls != bg && cls >= 0 && cls < C && n_gt > 0) {  // evaluators.py:167-168 otherwise
    const float ngt_f = (float)n_gt;
    const int rb = rank_base[b], tb = tp_base[b];
    const u64* in = keys + off[b];
    auto add_term = [&](const int r, const int cum) {  // r: 0-based rank in the class, cum: TPs up to and including
      if (r >= 1) {
        const float r0 = __fdiv_rn((float)(cum - 1), ngt_f), r1 = __fdiv_rn((float)cum, ngt_f);
        const double p0 = (double)(float)(cum - 1) / (double)r, p1 = (double)(float)cum / (double)(r + 1);
        acc += (double)__fsub_rn(r1, r0) * (p1 + p0) / 2.0;
      }
    };
    if (m <= kBucketCap) {
      // ---- key range of the bucket, 256 equal slices of it ----
      if (tid == 0) {
        s_min = kPad;
        s_max = 0ull;
      }
      for (int k = tid; k <= kSlices; k += kApThreads) {
        s_scnt[k] = 0;
        s_stp[k] = 0;
      }
      __syncthreads();
      u64 kmin = kPad, kmax = 0ull;
      for (int j = tid; j < m; j += kApThreads) {
        const u64 k = in[j];
        s_in[j] = k;
        kmin = k < kmin ? k : kmin;
        kmax = k > kmax ? k : kmax;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const u64 a = __shfl_xor_sync(0xFFFFFFFFu, kmin, o), c = __shfl_xor_sync(0xFFFFFFFFu, kmax, o);
        kmin = a < kmin ? a : kmin;
        kmax = c > kmax ? c : kmax;
      }
      if (lane == 0) {
        atomicMin(&s_min, kmin);
        atomicMax(&s_max, kmax);
      }
      __syncthreads();
      const u64 lo = s_min;
      const u64 span = s_max - lo;
      int sh = 0;
      while ((span >> sh) >= (1ull << 24)) ++sh;  // (key - lo) >> sh < 2^24
      const unsigned den = (unsigned)(span >> sh) + 1u;
      auto slice_of = [&](const u64 k) { return (int)(((unsigned)((k - lo) >> sh) * (unsigned)kSlices) / den); };
      for (int j = tid; j < m; j += kApThreads) {
        const u64 k = s_in[j];
        const int f = slice_of(k);
        atomicAdd(&s_scnt[f], 1);
        if (k & 1ull) atomicAdd(&s_stp[f], 1);
      }
      __syncthreads();
      if (warp == 0) {  // exclusive prefixes of the 256 slice counts and TP counts (8 per lane)
        int c8[8], t8[8], cs = 0, ts = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          c8[q] = s_scnt[lane * 8 + q];
          t8[q] = s_stp[lane * 8 + q];
          cs += c8[q];
          ts += t8[q];
        }
        int xc = cs, xt = ts;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int yc = __shfl_up_sync(0xFFFFFFFFu, xc, o), yt = __shfl_up_sync(0xFFFFFFFFu, xt, o);
          if (lane >= o) {
            xc += yc;
            xt += yt;
          }
        }
        int rc = xc - cs, rt = xt - ts;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          s_scnt[lane * 8 + q] = rc;
          s_stp[lane * 8 + q] = rt;
          s_cur[lane * 8 + q] = rc;
          rc += c8[q];
          rt += t8[q];
        }
        if (lane == 31) {
          s_scnt[kSlices] = rc;
          s_stp[kSlices] = rt;
        }
      }
      __syncthreads();
      for (int j = tid; j < m; j += kApThreads) {
        const u64 k = s_in[j];
        s_by[atomicAdd(&s_cur[slice_of(k)], 1)] = k;
      }
      __syncthreads();
      // ---- every TP key: keys / TP keys below it in its slice ----
      for (int j = tid; j < m; j += kApThreads) {
        const u64 k = s_by[j];
        if ((k & 1ull) == 0ull) continue;
        const int f = slice_of(k);
        const int f0 = s_scnt[f], f1 = s_scnt[f + 1];
        int below = 0, tp_below = 0;
        for (int q = f0; q < f1; ++q) {
          const u64 o = s_by[q];
          const int lt = o < k;
          below += lt;
          tp_below += lt & (int)(o & 1ull);
        }
        add_term(rb + f0 + below, tb + s_stp[f] + tp_below + 1);
      }
    } else {
      // ---- the sample was unlucky: count over the whole bucket in global memory (slow, exact) ----
      for (int j = tid; j < m; j += kApThreads) {
        const u64 k = in[j];
        if ((k & 1ull) == 0ull) continue;
        int below = 0, tp_below = 0;
        for (int q = 0; q < m; ++q) {
          const u64 o = in[q];
          const int lt = o < k;
          below += lt;
          tp_below += lt & (int)(o & 1ull);
        }
        add_term(rb + below, tb + tp_below + 1);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
  if (lane == 0) s_red[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kApThreads / 32; ++w) t += s_red[w];
    partial[b] = t;
  }
}

// F: per class (one warp each), the partials of its buckets: lane l adds buckets l, l + 32, ... in order, then a fixed
// shuffle tree - the same order every run
__global__ void __launch_bounds__(32)
    ssdx_ap_sum_kernel(const int nb, const int* __restrict__ bcls, const double* __restrict__ partial,
                       double* __restrict__ out_ap) {
  const int c = blockIdx.x, lane = threadIdx.x;
  double t = 0.0;
  for (int b = lane; b < nb; b += 32)
    if (bcls[b] == c) t += partial[b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
  if (lane == 0) out_ap[c] = t;
}

__global__ void ssdx_ap_zero_kernel(double* __restrict__ out_ap, const int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out_ap[c] = 0.0;
}

struct ApWorkspace {
  size_t keys, bid, splitters, ints, partial, total;
};
inline ApWorkspace ap_workspace(long long n, int C) {
  (void)C;
  ApWorkspace w;
  size_t off = 0;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  w.keys = off;      off += ssdx_align_up(nn * 8, 256);
  w.bid = off;       off += ssdx_align_up(nn * 2, 256);
  w.splitters = off; off += ssdx_align_up((size_t)kMaxSplit * 8, 256);
  // tpc | bcls | cnt | off | rank_base | tp_base (kMaxBuckets ints each) | cta_cnt (kApCtas x kMaxBuckets)
  w.ints = off;      off += ssdx_align_up((size_t)(6 + kApCtas) * kMaxBuckets * 4, 256);
  w.partial = off;   off += ssdx_align_up((size_t)kMaxBuckets * 8, 256);
  w.total = off;
  return w;
}

}  // namespace

extern "C" size_t ssdx_average_precision_workspace_bytes(int64_t num_predictions, int32_t num_classes) {
  if (num_predictions < 0 || num_predictions >= (1ll << 24) || num_classes < 1 || num_classes > 254) return 0;
  return ap_workspace(num_predictions, num_classes).total;
}

extern "C" int ssdx_average_precision(const uint8_t* tp, const float* confidence, const int32_t* label,
                                      int64_t num_predictions, int32_t num_classes, const int64_t* counts,
                                      int32_t background_label, double* out_ap, void* workspace,
                                      size_t workspace_bytes, int device, ssdx_stream stream) {
  SSDX_CHECK_ARG(num_classes >= 1 && num_classes <= 254, "ssdx_average_precision: num_classes must be in [1, 254]");
  SSDX_CHECK_ARG(num_predictions >= 0 && num_predictions < (1ll << 24),
                 "ssdx_average_precision: at most 2^24 - 1 predictions per call (got %lld)",
                 (long long)num_predictions);
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
  const int C = num_classes;
  if (num_predictions == 0) {
    ssdx_ap_zero_kernel<<<(C + 127) / 128, 128, 0, s>>>(out_ap, C);
    SSDX_LAUNCH_CHECK("ssdx_ap_zero_kernel");
    return SSDX_OK;
  }
  char* base = (char*)workspace;
  u64* keys = (u64*)(base + w.keys);
  unsigned short* bid = (unsigned short*)(base + w.bid);
  u64* splitters = (u64*)(base + w.splitters);
  int* tpc = (int*)(base + w.ints);
  int* bcls = tpc + kMaxBuckets;
  int* cnt = bcls + kMaxBuckets;
  int* off = cnt + kMaxBuckets;
  int* rank_base = off + kMaxBuckets;
  int* tp_base = rank_base + kMaxBuckets;
  int* cta_cnt = tp_base + kMaxBuckets;
  double* partial = (double*)(base + w.partial);
  const long long n = num_predictions;
  int S = (int)((n + kBucketTarget - 1) / kBucketTarget);  // sample buckets; <= n, so the sample has >= S entries
  S = S < 1 ? 1 : (S > kMaxSplit ? kMaxSplit : S);
  const int nb = S + C;  // bucket ids: (splitters below) + class
  SSDX_CUDA(cudaMemsetAsync(tpc, 0, (size_t)kMaxBuckets * 4, s));
  SSDX_CUDA(cudaMemsetAsync(bcls, 0xFF, (size_t)kMaxBuckets * 4, s));  // -1: empty bucket
  ssdx_ap_sample_kernel<<<1, 1024, 0, s>>>(confidence, label, n, C, background_label, S, splitters);
  SSDX_LAUNCH_CHECK("ssdx_ap_sample_kernel");
  const int n_ctas = (int)(n < (long long)kApCtas * 1024 ? (n + 1023) / 1024 : kApCtas);
  ssdx_ap_bucket_kernel<<<n_ctas, 1024, 0, s>>>(tp, confidence, label, n, C, background_label, splitters, S - 1, nb, bid,
                                               cta_cnt, tpc, bcls);
  SSDX_LAUNCH_CHECK("ssdx_ap_bucket_kernel");
  ssdx_ap_columns_kernel<<<(nb + 7) / 8, 256, 0, s>>>(cta_cnt, n_ctas, nb, cnt);
  SSDX_LAUNCH_CHECK("ssdx_ap_columns_kernel");
  ssdx_ap_offsets_kernel<<<1, 1024, 0, s>>>(nb, tpc, bcls, cnt, off, rank_base, tp_base);
  SSDX_LAUNCH_CHECK("ssdx_ap_offsets_kernel");
  ssdx_ap_scatter_kernel<<<n_ctas, 512, 0, s>>>(tp, confidence, bid, n, nb, off, cta_cnt, keys);
  SSDX_LAUNCH_CHECK("ssdx_ap_scatter_kernel");
  const size_t smem = (size_t)2 * kBucketCap * sizeof(u64);
  SSDX_CUDA(cudaFuncSetAttribute(ssdx_ap_terms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ssdx_ap_terms_kernel<<<nb, kApThreads, smem, s>>>(keys, cnt, off, rank_base, tp_base, bcls, (const long long*)counts,
                                                    C, background_label, partial);
  SSDX_LAUNCH_CHECK("ssdx_ap_terms_kernel");
  ssdx_ap_sum_kernel<<<C, 32, 0, s>>>(nb, bcls, partial, out_ap);
  SSDX_LAUNCH_CHECK("ssdx_ap_sum_kernel");
  return SSDX_OK;
}
