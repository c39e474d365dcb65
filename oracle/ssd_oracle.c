/*
 * ssd_oracle.c - plain C (+OpenMP over images) restatement of the reference's SSD box path.
 *
 * TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT: used by tests/ (cross-check against oracle/ssd_oracle.py)
 * and by bench.py's cpu_baseline / `--impl reference` legs as the timed host-CPU port of the reference
 * (the reference itself needs TensorFlow 2.13, which is not installable here).  The shipped library
 * never links or calls this file.
 *
 * Follows, statement by statement:
 *   encode : ssdseglib/datacoder.py:205-300 (IoU matrix materialised per image, as the reference does;
 *            arg-max ties -> lowest index; concat -> unique -> sequential scatter, last writer wins)
 *   decode : ssdseglib/layers.py:57-79 (batched yxyx) and datacoder.py:367-432 (masked)
 *   nms    : ssdseglib/layers.py:141-162 + TF CombinedNonMaxSuppression semantics
 *   eval   : ssdseglib/evaluators.py:27-63, :129-153
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC (no -march=native: no FMA).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

static inline float fmax2(float a, float b) { return a > b ? a : b; }
static inline float fmin2(float a, float b) { return a < b ? a : b; }

/* ------------------------------------------------------------------ encode */
static void encode_one(const float* anc, int A, const float* gt, const int32_t* glab, int G, int C, float thr,
                       const float* sd, float* labels, float* offsets, int32_t* match, float* iou /* A*G scratch */,
                       int32_t* rows /* 2*(A+G) scratch */) {
  for (int a = 0; a < A; ++a) {
    for (int c = 0; c < C; ++c) labels[(size_t)a * C + c] = (c == 0) ? 1.0f : 0.0f; /* :274-283 */
    offsets[4 * a + 0] = offsets[4 * a + 1] = offsets[4 * a + 2] = offsets[4 * a + 3] = 0.0f;
    if (match) match[a] = -1;
  }
  if (G <= 0) return;
  /* IoU matrix, :205-220 */
  for (int a = 0; a < A; ++a) {
    const float ax0 = anc[4 * a], ay0 = anc[4 * a + 1], ax1 = anc[4 * a + 2], ay1 = anc[4 * a + 3];
    const float area_a = (ay1 - ay0 + 1.0f) * (ax1 - ax0 + 1.0f);
    for (int g = 0; g < G; ++g) {
      const float gx0 = gt[4 * g], gy0 = gt[4 * g + 1], gx1 = gt[4 * g + 2], gy1 = gt[4 * g + 3];
      const float area_g = (gx1 - gx0 + 1.0f) * (gy1 - gy0 + 1.0f);
      const float iw = fmax2(0.0f, fmin2(ax1, gx1) - fmax2(ax0, gx0) + 1.0f);
      const float ih = fmax2(0.0f, fmin2(ay1, gy1) - fmax2(ay0, gy0) + 1.0f);
      const float inter = iw * ih;
      iou[(size_t)a * G + g] = inter / (area_a + area_g - inter);
    }
  }
  /* step 1 (:230-231) and step 2 (:236-241) -> concatenated (anchor, gt) rows */
  int n = 0;
  for (int g = 0; g < G; ++g) {
    int best = 0;
    float bv = iou[g];
    for (int a = 1; a < A; ++a)
      if (iou[(size_t)a * G + g] > bv) { bv = iou[(size_t)a * G + g]; best = a; }
    if (bv > 0.0f) { rows[2 * n] = best; rows[2 * n + 1] = g; ++n; }
  }
  const int n1 = n;
  for (int a = 0; a < A; ++a) {
    int best = 0;
    float bv = iou[(size_t)a * G];
    for (int g = 1; g < G; ++g)
      if (iou[(size_t)a * G + g] > bv) { bv = iou[(size_t)a * G + g]; best = g; }
    if (bv > thr) {
      /* UniqueV2 (:244): drop the row if the same (anchor, gt) already came from step 1 */
      int dup = 0;
      for (int r = 0; r < n1; ++r)
        if (rows[2 * r] == a && rows[2 * r + 1] == best) { dup = 1; break; }
      if (!dup) { rows[2 * n] = a; rows[2 * n + 1] = best; ++n; }
    }
  }
  /* gather, offsets (:246-269), sequential scatter (:286-298): last writer wins */
  for (int r = 0; r < n; ++r) {
    const int a = rows[2 * r], g = rows[2 * r + 1];
    const float ax0 = anc[4 * a], ay0 = anc[4 * a + 1], ax1 = anc[4 * a + 2], ay1 = anc[4 * a + 3];
    const float gx0 = gt[4 * g], gy0 = gt[4 * g + 1], gx1 = gt[4 * g + 2], gy1 = gt[4 * g + 3];
    const float dcx = (ax1 + ax0) / 2.0f, dcy = (ay1 + ay0) / 2.0f, dw = ax1 - ax0 + 1.0f, dh = ay1 - ay0 + 1.0f;
    const float gcx = (gx1 + gx0) / 2.0f, gcy = (gy1 + gy0) / 2.0f, gw = gx1 - gx0 + 1.0f, gh = gy1 - gy0 + 1.0f;
    offsets[4 * a + 0] = (gcx - dcx) / dw / sd[0];
    offsets[4 * a + 1] = (gcy - dcy) / dh / sd[1];
    offsets[4 * a + 2] = logf(gw / dw + 1.0f) / sd[2];
    offsets[4 * a + 3] = logf(gh / dh + 1.0f) / sd[3];
    for (int c = 0; c < C; ++c) labels[(size_t)a * C + c] = (glab[g] == c) ? 1.0f : 0.0f;
    if (match) match[a] = g;
  }
}

EXPORT void oracle_encode_batch(const float* anc, int A, const float* gt, const int32_t* glab, const int32_t* gcnt,
                                int B, int Gmax, int C, float thr, const float* sd, float* labels, float* offsets,
                                int32_t* match, int nthreads) {
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    float* iou = (float*)malloc(sizeof(float) * (size_t)A * (Gmax > 0 ? Gmax : 1));
    int32_t* rows = (int32_t*)malloc(sizeof(int32_t) * 2 * ((size_t)A + Gmax + 1));
#pragma omp for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b)
      encode_one(anc, A, gt + (size_t)b * Gmax * 4, glab + (size_t)b * Gmax, gcnt[b], C, thr, sd,
                 labels + (size_t)b * A * C, offsets + (size_t)b * A * 4, match ? match + (size_t)b * A : NULL, iou,
                 rows);
    free(iou);
    free(rows);
  }
}

/* ------------------------------------------------------------------ decode */
EXPORT void oracle_decode_batch(const float* cent /*A,4 cx cy w h*/, int A, const float* off, int B, const float* sd,
                                int mode /*0 yxyx, 1 centroids masked, 2 corners masked*/, float* out, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
  for (int b = 0; b < B; ++b) {
    for (int a = 0; a < A; ++a) {
      const float* t = off + ((size_t)b * A + a) * 4;
      float* o = out + ((size_t)b * A + a) * 4;
      const float dcx = cent[4 * a], dcy = cent[4 * a + 1], dw = cent[4 * a + 2], dh = cent[4 * a + 3];
      float cx = t[0] * sd[0] * dw + dcx;
      float cy = t[1] * sd[1] * dh + dcy;
      float w = (expf(t[2] * sd[2]) - 1.0f) * dw;
      float h = (expf(t[3] * sd[3]) - 1.0f) * dh;
      if (mode != 0) {
        const float s = fabsf(t[0]) + fabsf(t[1]) + fabsf(t[2]) + fabsf(t[3]);
        const float nb = s > 0.0f ? 1.0f : 0.0f;
        cx *= nb; cy *= nb; w *= nb; h *= nb;
        if (mode == 1) { o[0] = cx; o[1] = cy; o[2] = w; o[3] = h; continue; }
      }
      const float x0 = cx - (w - 1.0f) / 2.0f, y0 = cy - (h - 1.0f) / 2.0f;
      const float x1 = cx + (w - 1.0f) / 2.0f, y1 = cy + (h - 1.0f) / 2.0f;
      if (mode == 0) { o[0] = y0; o[1] = x0; o[2] = y1; o[3] = x1; continue; }
      const float s2 = fabsf(cx) + fabsf(cy) + fabsf(w) + fabsf(h);
      const float nb2 = s2 > 0.0f ? 1.0f : 0.0f;
      o[0] = x0 * nb2; o[1] = y0 * nb2; o[2] = x1 * nb2; o[3] = y1 * nb2;
    }
  }
}

/* ------------------------------------------------------------------ NMS */
typedef struct { float score; int32_t idx; int32_t cls; } cand_t;

static int cand_cmp(const void* pa, const void* pb) {
  const cand_t* a = (const cand_t*)pa; const cand_t* b = (const cand_t*)pb;
  if (a->score > b->score) return -1;
  if (a->score < b->score) return 1;
  if (a->cls != b->cls) return a->cls < b->cls ? -1 : 1;
  return a->idx < b->idx ? -1 : (a->idx > b->idx ? 1 : 0);
}

static float nms_iou(const float* bi, const float* bj) {
  const float y0i = fmin2(bi[0], bi[2]), x0i = fmin2(bi[1], bi[3]), y1i = fmax2(bi[0], bi[2]), x1i = fmax2(bi[1], bi[3]);
  const float y0j = fmin2(bj[0], bj[2]), x0j = fmin2(bj[1], bj[3]), y1j = fmax2(bj[0], bj[2]), x1j = fmax2(bj[1], bj[3]);
  const float ai = (y1i - y0i) * (x1i - x0i), aj = (y1j - y0j) * (x1j - x0j);
  if (ai <= 0 || aj <= 0) return 0.0f;
  const float ih = fmax2(fmin2(y1i, y1j) - fmax2(y0i, y0j), 0.0f);
  const float iw = fmax2(fmin2(x1i, x1j) - fmax2(x0i, x0j), 0.0f);
  const float inter = ih * iw;
  return inter / (ai + aj - inter);
}

EXPORT void oracle_combined_nms(const float* boxes /*B,A,4 yxyx*/, const float* scores /*B,A,C*/, int B, int A, int C,
                                int max_per_class, int max_total, float iou_thr, float score_thr,
                                float* out /*B,K,6*/, int32_t* valid /*B*/, int32_t* aidx /*B,K*/, int nthreads) {
  const int limit = max_per_class < A ? max_per_class : A;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
  {
    cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * (size_t)A);
    cand_t* picks = (cand_t*)malloc(sizeof(cand_t) * (size_t)C * limit);
    int32_t* sel = (int32_t*)malloc(sizeof(int32_t) * (size_t)limit);
#pragma omp for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
      const float* bx = boxes + (size_t)b * A * 4;
      int np = 0;
      for (int c = 0; c < C; ++c) {
        int n = 0;
        for (int i = 0; i < A; ++i) {
          const float s = scores[((size_t)b * A + i) * C + c];
          if (s > score_thr) { cand[n].score = s; cand[n].idx = i; cand[n].cls = c; ++n; }
        }
        qsort(cand, (size_t)n, sizeof(cand_t), cand_cmp);
        int ns = 0;
        for (int k = 0; k < n && ns < limit; ++k) {
          int ok = 1;
          for (int j = ns - 1; j >= 0; --j)
            if (nms_iou(bx + 4 * (size_t)cand[k].idx, bx + 4 * (size_t)sel[j]) > iou_thr) { ok = 0; break; }
          if (ok) { sel[ns++] = cand[k].idx; picks[np++] = cand[k]; }
        }
      }
      qsort(picks, (size_t)np, sizeof(cand_t), cand_cmp);
      const int v = np < max_total ? np : max_total;
      if (valid) valid[b] = v;
      float* o = out + (size_t)b * max_total * 6;
      memset(o, 0, sizeof(float) * (size_t)max_total * 6);
      for (int k = 0; k < max_total; ++k) {
        if (aidx) aidx[(size_t)b * max_total + k] = k < v ? picks[k].idx : -1;
        if (k >= v) continue;
        const float* p = bx + 4 * (size_t)picks[k].idx;
        o[6 * k + 0] = (float)picks[k].cls; o[6 * k + 1] = picks[k].score;
        o[6 * k + 2] = p[1]; o[6 * k + 3] = p[0]; o[6 * k + 4] = p[3]; o[6 * k + 5] = p[2];
      }
    }
    free(cand); free(picks); free(sel);
  }
}

/* ------------------------------------------------------------------ eval */
EXPORT void oracle_eval_tp(const int32_t* plab /*B,K*/, const float* pbox /*B,K,4*/, const float* gbox /*B,G,4*/,
                           const int32_t* glab, const int32_t* gcnt, int B, int K, int Gmax, int C, int bg, float thr,
                           uint8_t* tp /*B,K*/, int64_t* counts /*C,3*/, int nthreads) {
  /* images are independent (evaluators.py:106); the integer counts are summed per thread and added at the end */
  if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
  {
    int64_t* local = (int64_t*)calloc((size_t)3 * C, sizeof(int64_t));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int b = 0; b < B; ++b) {
      const int G = gcnt[b];
      for (int g = 0; g < G; ++g) {
        const int l = glab[(size_t)b * Gmax + g];
        if (l >= 0 && l < C) local[3 * l + 2] += 1;
      }
      for (int k = 0; k < K; ++k) {
        const int pl = plab[(size_t)b * K + k];
        uint8_t flag = 0;
        if (pl != bg) {
          const float* p = pbox + ((size_t)b * K + k) * 4;
          const float pa = (p[2] - p[0] + 1.0f) * (p[3] - p[1] + 1.0f);
          float best = (G == 0) ? 0.0f : -INFINITY;
          for (int g = 0; g < G; ++g) {
            const float* t = gbox + ((size_t)b * Gmax + g) * 4;
            const float ta = (t[2] - t[0] + 1.0f) * (t[3] - t[1] + 1.0f);
            const float iw = fmax2(0.0f, fmin2(p[2], t[2]) - fmax2(p[0], t[0]) + 1.0f);
            const float ih = fmax2(0.0f, fmin2(p[3], t[3]) - fmax2(p[1], t[1]) + 1.0f);
            const float inter = iw * ih;
            float iou = inter / (pa + ta - inter + 1e-7f);
            iou = iou * ((glab[(size_t)b * Gmax + g] == pl) ? 1.0f : 0.0f);
            if (iou > best) best = iou;
          }
          flag = best >= thr;
          if (pl >= 0 && pl < C) local[3 * pl + (flag ? 0 : 1)] += 1;
        }
        tp[(size_t)b * K + k] = flag;
      }
    }
#ifdef _OPENMP
#pragma omp critical
#endif
    for (int i = 0; i < 3 * C; ++i) counts[i] += local[i];
    free(local);
  }
}
