// libssdx: library-level entry points, the anchor table and the host-side GT ingest.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <math.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "grid_window.cuh"

static thread_local char g_err[512] = "";

void ssdx_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool ssdx_debug_sync() {
  static const bool on = [] {  // read once: the check sits behind every launch
    const char* e = getenv("SSDX_DEBUG_SYNC");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

static std::atomic<int> g_encode_path(SSDX_ENC_AUTO);
static std::atomic<int> g_nms_path(0);
static std::atomic<int> g_nms_cta(0);
int ssdx_option_nms_cta() { return g_nms_cta.load(std::memory_order_relaxed); }
int ssdx_option_encode_path() { return g_encode_path.load(std::memory_order_relaxed); }
const char* ssdx_option_nms_path() {
  static const char* const names[5] = {nullptr, "perclass", "pcta", "generic", "blocked"};
  const int i = g_nms_path.load(std::memory_order_relaxed);
  return names[(i >= 0 && i < 5) ? i : 0];
}

int ssdx_sm_count(int device) {
  static int cache[64] = {0};
  if (device < 0 || device >= 64) return 148;
  if (cache[device] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
    cache[device] = n;
  }
  return cache[device];
}

// ---------------------------------------------------------------------------------------------
// Grid detection (encode_grid.cu).  boxes.DefaultBoundingBoxes lays the anchors out map-major, then row, column,
// box shape (boxes.py:198-205), with corners = centre -/+ (size - 1) / 2 from a meshgrid of per-axis centres
// (boxes.py:112-146) - so inside one map the x extents of (row, col, k) do not depend on the row and the y extents do
// not depend on the column, bit for bit, whatever rescale or corner/centroid conversion came after.  The table is
// scanned for that structure; any table that does not have it (subsets, permutations, hand-made anchors) keeps the
// general tile match.  Only exact equalities decide; the centre model below merely sizes windows and carries its own
// measured error.
// ---------------------------------------------------------------------------------------------
namespace {

struct GridMap { size_t off; int H, W, nb; };

inline bool same_x(const float* c, size_t i, size_t j) {
  return memcmp(c + 4 * i, c + 4 * j, 4) == 0 && memcmp(c + 4 * i + 2, c + 4 * j + 2, 4) == 0;
}
inline bool same_y(const float* c, size_t i, size_t j) {
  return memcmp(c + 4 * i + 1, c + 4 * j + 1, 4) == 0 && memcmp(c + 4 * i + 3, c + 4 * j + 3, 4) == 0;
}

// rows [r_from, r_to) of the candidate map repeat the x extents of row 0 and hold one y extent per shape
bool rows_ok(const float* c, size_t off, int r_from, int r_to, int W, int nb) {
  const size_t L = (size_t)W * nb;
  for (int r = r_from; r < r_to; ++r) {
    const size_t row = off + (size_t)r * L;
    for (size_t j = 0; j < L; ++j) {
      if (!same_x(c, off + j, row + j)) return false;
      if (!same_y(c, row + j % nb, row + j)) return false;
    }
  }
  return true;
}

// model + pads of one (map, shape); false = unusable for windows (the caller tries another decomposition)
bool make_combo(const float* c, const float* area, const GridMap& m, int k, ssdx_grid_combo* out) {
  const int H = m.H, W = m.W, nb = m.nb;
  const size_t L = (size_t)W * nb;
  auto X0 = [&](int col) { return (double)c[4 * (m.off + (size_t)col * nb + k) + 0]; };
  auto X1 = [&](int col) { return (double)c[4 * (m.off + (size_t)col * nb + k) + 2]; };
  auto Y0 = [&](int row) { return (double)c[4 * (m.off + (size_t)row * L + k) + 1]; };
  auto Y1 = [&](int row) { return (double)c[4 * (m.off + (size_t)row * L + k) + 3]; };
  double hx = -1e300, hy = -1e300, maxabs = 0;
  for (int col = 0; col < W; ++col) {
    if (!(X1(col) - X0(col) + 1.0 > 0.0)) return false;  // (boxes smaller than one cell have x1 < x0: boxes.py:143)
    hx = std::max(hx, 0.5 * (X1(col) - X0(col)));
    maxabs = std::max(maxabs, std::max(fabs(X0(col)), fabs(X1(col))));
  }
  for (int row = 0; row < H; ++row) {
    if (!(Y1(row) - Y0(row) + 1.0 > 0.0)) return false;
    hy = std::max(hy, 0.5 * (Y1(row) - Y0(row)));
    maxabs = std::max(maxabs, std::max(fabs(Y0(row)), fabs(Y1(row))));
  }
  const double cx0 = 0.5 * (X0(0) + X1(0)), cy0 = 0.5 * (Y0(0) + Y1(0));
  const double dx = W > 1 ? (0.5 * (X0(W - 1) + X1(W - 1)) - cx0) / (W - 1) : 1.0;
  const double dy = H > 1 ? (0.5 * (Y0(H - 1) + Y1(H - 1)) - cy0) / (H - 1) : 1.0;
  if (!(dx > 0.0) || !(dy > 0.0)) return false;
  double ex = 0, ey = 0;
  for (int col = 0; col < W; ++col)
    ex = std::max(ex, fabs(0.5 * (X0(col) + X1(col)) - (cx0 + col * dx)) + fabs(0.5 * (X1(col) - X0(col)) - hx));
  for (int row = 0; row < H; ++row)
    ey = std::max(ey, fabs(0.5 * (Y0(row) + Y1(row)) - (cy0 + row * dy)) + fabs(0.5 * (Y1(row) - Y0(row)) - hy));
  if ((W > 1 && ex > 0.25 * dx) || (H > 1 && ey > 0.25 * dy)) return false;  // not a regular grid after all
  float amin = INFINITY;
  for (int row = 0; row < H; ++row)
    for (int col = 0; col < W; ++col) {
      const float a = area[m.off + ((size_t)row * W + col) * nb + k];
      if (!(a > 0.0f) || !std::isfinite(a)) return false;
      amin = std::min(amin, a);
    }
  // rounding margin: the device evaluates the window bounds in fp32 on coordinates of this magnitude
  const double margin = 0.01 + 1e-5 * maxabs;
  out->off = (int)(m.off + k);
  out->H = H; out->W = W; out->nb = nb;
  out->cx0 = (float)cx0; out->inv_dx = (float)(1.0 / dx);
  out->cy0 = (float)cy0; out->inv_dy = (float)(1.0 / dy);
  out->hx = (float)(hx + 1e-6 * fabs(hx)); out->hy = (float)(hy + 1e-6 * fabs(hy));
  out->px = (float)(1.01 * ex + margin); out->py = (float)(1.01 * ey + margin);
  out->aw = (float)((2 * hx + 1) * (1 + 1e-6)); out->ah = (float)((2 * hy + 1) * (1 + 1e-6));
  out->amin = amin;
  out->ax = 0;  // set by detect_grids
  return std::isfinite(out->inv_dx) && std::isfinite(out->inv_dy) && std::isfinite(out->px) && std::isfinite(out->py);
}

bool detect_grids(const float* c, const float* area, size_t A, std::vector<ssdx_grid_combo>& combos,
                  std::vector<ssdx_grid_span>& spans, std::vector<float>& axes,
                  std::vector<int>* combo_span = nullptr) {
  combos.clear();
  spans.clear();
  axes.clear();
  if (combo_span) combo_span->clear();
  for (size_t i = 0; i < 4 * A; ++i)
    if (!std::isfinite(c[i])) return false;
  size_t off = 0;
  while (off < A) {
    const size_t rem = A - off;
    bool found = false;
    for (int nb = 1; nb <= 64 && (size_t)nb <= rem && !found; ++nb) {
      int W = 0, H = 0;
      if ((size_t)nb == rem) {
        W = 1; H = 1;  // one cell at the end of the table (a 1 x 1 feature map)
      } else if (same_y(c, off, off + nb)) {  // next column
        for (size_t col = 2; col * nb <= rem; ++col) {
          const size_t i = off + col * nb;
          if (i == A) { W = (int)col; H = 1; break; }                        // one-row map ends the table
          if (!same_y(c, off, i)) { W = (int)col; if (!same_x(c, off, i)) H = 1; break; }  // next row, or next map
          if (col > 65535) break;
        }
      } else if (same_x(c, off, off + nb)) {
        W = 1;  // one column: the next cell is the next row
      }
      if (W == 0 || (size_t)W * nb > 16384) continue;
      const size_t L = (size_t)W * nb;
      if (H == 0) {
        H = 1;
        while (off + (size_t)(H + 1) * L <= A && H < 32767 && rows_ok(c, off, H, H + 1, W, nb)) ++H;
      }
      if (!rows_ok(c, off, 0, 1, W, nb)) continue;  // row 0 itself: one y extent per shape
      const GridMap m{off, H, W, nb};
      std::vector<ssdx_grid_combo> mine(nb);
      bool ok = combos.size() + nb <= (size_t)kSsdxMaxCombos;
      for (int k = 0; k < nb && ok; ++k) ok = make_combo(c, area, m, k, &mine[k]);
      if (!ok) continue;
      const int combo0 = (int)combos.size();
      for (int k = 0; k < nb; ++k) {  // axis arrays: x0[W] | x1[W] | y0[H] | y1[H], the floats of the table
        mine[k].ax = (int)axes.size();
        for (int col = 0; col < W; ++col) axes.push_back(c[4 * (off + (size_t)col * nb + k) + 0]);
        for (int col = 0; col < W; ++col) axes.push_back(c[4 * (off + (size_t)col * nb + k) + 2]);
        for (int row = 0; row < H; ++row) axes.push_back(c[4 * (off + (size_t)row * L + k) + 1]);
        for (int row = 0; row < H; ++row) axes.push_back(c[4 * (off + (size_t)row * L + k) + 3]);
        while (axes.size() % 4) axes.push_back(0.0f);  // 16-byte granules for the vector copy into shared memory
      }
      combos.insert(combos.end(), mine.begin(), mine.end());
      const int rps = std::max(1, (int)(kSsdxSpanAnchors / L));
      if (combo_span)
        for (int k = 0; k < nb; ++k) {
          combo_span->push_back((int)spans.size());
          combo_span->push_back(rps);
        }
      for (int r = 0; r < H; r += rps) {
        ssdx_grid_span s;
        s.row0 = r; s.nrows = std::min(rps, H - r);
        s.a_lo = (int)(off + (size_t)r * L); s.n = (int)((size_t)s.nrows * L);
        s.combo0 = combo0; s.nb = nb; s.pad0 = s.pad1 = 0;
        spans.push_back(s);
      }
      off += (size_t)H * L;
      found = true;
    }
    if (!found) {
      combos.clear(); spans.clear(); axes.clear();
      if (combo_span) combo_span->clear();
      return false;
    }
  }
  return !combos.empty();
}

}  // namespace

extern "C" {

int ssdx_version(void) { return SSDX_VERSION; }

const char* ssdx_last_error(void) { return g_err; }

int ssdx_set_option(const char* key, const char* value) {
  SSDX_CHECK_ARG(key != nullptr && value != nullptr, "ssdx_set_option: NULL argument");
  if (strcmp(key, "encode_path") == 0) {
    static const char* const v[5] = {"auto", "tile_fused", "tile_tiled", "grid_fused", "grid_split"};
    for (int i = 0; i < 5; ++i)
      if (strcmp(value, v[i]) == 0) {
        g_encode_path.store(i);
        return SSDX_OK;
      }
  } else if (strcmp(key, "nms_path") == 0) {
    static const char* const v[5] = {"auto", "perclass", "pcta", "generic", "blocked"};
    for (int i = 0; i < 5; ++i)
      if (strcmp(value, v[i]) == 0) {
        g_nms_path.store(i);
        return SSDX_OK;
      }
  } else if (strcmp(key, "nms_cta") == 0) {
    static const char* const v[4] = {"auto", "256", "512", "1024"};
    static const int w[4] = {0, 256, 512, 1024};
    for (int i = 0; i < 4; ++i)
      if (strcmp(value, v[i]) == 0) {
        g_nms_cta.store(w[i]);
        return SSDX_OK;
      }
  } else {
    ssdx_set_error("ssdx_set_option: unknown key '%s'", key);
    return SSDX_ERR_INVALID;
  }
  ssdx_set_error("ssdx_set_option: bad value '%s' for '%s'", value, key);
  return SSDX_ERR_INVALID;
}

int ssdx_device_count(int* count) {
  SSDX_CHECK_ARG(count != nullptr, "ssdx_device_count: count is NULL");
  *count = 0;
  SSDX_CUDA(cudaGetDeviceCount(count));
  return SSDX_OK;
}

int ssdx_anchors_create(const float* corners_xyxy_host, const float* centroids_host, int32_t num_anchors, int device,
                        ssdx_anchors** out) {
  SSDX_CHECK_ARG(out != nullptr, "ssdx_anchors_create: out is NULL");
  *out = nullptr;
  SSDX_CHECK_ARG(corners_xyxy_host != nullptr, "you must pass all default bounding boxes corners coordinates!");
  SSDX_CHECK_ARG(num_anchors > 0, "ssdx_anchors_create: num_anchors must be > 0 (got %d)", num_anchors);
  SSDX_DEVICE(device);
  const size_t A = (size_t)num_anchors;
  std::vector<float> area(A), enc(4 * A), dec(4 * A);
  for (size_t a = 0; a < A; ++a) {
    // volatile: one fp32 rounding per operation on any host compiler setting
    volatile float x0 = corners_xyxy_host[4 * a + 0], y0 = corners_xyxy_host[4 * a + 1];
    volatile float x1 = corners_xyxy_host[4 * a + 2], y1 = corners_xyxy_host[4 * a + 3];
    volatile float hh = y1 - y0, ww = x1 - x0;
    hh = hh + 1.0f;
    ww = ww + 1.0f;
    volatile float ar = hh * ww;  // datacoder.py:112
    area[a] = ar;
    volatile float sx = x1 + x0, sy = y1 + y0;
    enc[4 * a + 0] = sx / 2.0f;  // datacoder.py:141-144
    enc[4 * a + 1] = sy / 2.0f;
    enc[4 * a + 2] = ww;
    enc[4 * a + 3] = hh;
  }
  if (centroids_host != nullptr)
    memcpy(dec.data(), centroids_host, 4 * A * sizeof(float));
  else
    dec = enc;
  // ---- match order (encode): sort by (area half-octave descending, Morton code of the centre), tile, bound ----
  const float inf = INFINITY;
  double cx_lo = 1e300, cx_hi = -1e300, cy_lo = 1e300, cy_hi = -1e300;
  for (size_t a = 0; a < A; ++a) {
    const double cx = enc[4 * a + 0], cy = enc[4 * a + 1];
    if (std::isfinite(cx)) { cx_lo = std::min(cx_lo, cx); cx_hi = std::max(cx_hi, cx); }
    if (std::isfinite(cy)) { cy_lo = std::min(cy_lo, cy); cy_hi = std::max(cy_hi, cy); }
  }
  auto spread = [](uint32_t v) {  // 16 bits -> even bit positions
    v &= 0xFFFFu;
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
  };
  auto quant = [](double v, double lo, double hi) -> uint32_t {
    if (!(hi > lo) || !std::isfinite(v)) return 0u;
    const double q = (v - lo) / (hi - lo) * 1023.0;
    return (uint32_t)std::min(1023.0, std::max(0.0, q));
  };
  std::vector<uint64_t> key(A);
  for (size_t a = 0; a < A; ++a) {
    const float ar = area[a];
    int bucket = 0;  // degenerate anchors (area <= 0, NaN) sort last; they never intersect anything
    if (ar > 0.0f && std::isfinite(ar)) bucket = 600 + (int)floor(2.0 * log2((double)ar));
    bucket = std::min(1023, std::max(0, bucket));
    const uint32_t m = spread(quant(enc[4 * a + 0], cx_lo, cx_hi)) | (spread(quant(enc[4 * a + 1], cy_lo, cy_hi)) << 1);
    key[a] = ((uint64_t)(1023 - bucket) << 52) | ((uint64_t)m << 32) | (uint64_t)a;  // index: total order
  }
  std::sort(key.begin(), key.end());
  const size_t n_tiles = (A + kSsdxTile - 1) / kSsdxTile, Ap = n_tiles * kSsdxTile;
  std::vector<float> m_box(4 * Ap), m_area(Ap, 1.0f), t_bbox(4 * n_tiles), t_area(2 * n_tiles);
  std::vector<int> m_idx(Ap, -1);
  for (size_t i = 0; i < Ap; ++i) {
    if (i < A) {
      const size_t a = (size_t)(key[i] & 0xFFFFFFFFull);
      memcpy(&m_box[4 * i], corners_xyxy_host + 4 * a, 4 * sizeof(float));
      m_area[i] = area[a];
      m_idx[i] = (int)a;
    } else {
      m_box[4 * i + 0] = inf; m_box[4 * i + 1] = inf; m_box[4 * i + 2] = -inf; m_box[4 * i + 3] = -inf;
    }
  }
  for (size_t t = 0; t < n_tiles; ++t) {
    float bx0 = inf, by0 = inf, bx1 = -inf, by1 = -inf, amin = inf, amax = -inf;
    bool odd = false;
    for (size_t i = t * kSsdxTile; i < (t + 1) * kSsdxTile && i < A; ++i) {
      bx0 = std::min(bx0, m_box[4 * i + 0]); by0 = std::min(by0, m_box[4 * i + 1]);
      bx1 = std::max(bx1, m_box[4 * i + 2]); by1 = std::max(by1, m_box[4 * i + 3]);
      amin = std::min(amin, m_area[i]); amax = std::max(amax, m_area[i]);
      odd = odd || !(m_area[i] > 0.0f) || !std::isfinite(m_area[i]) || !std::isfinite(m_box[4 * i + 0]) ||
            !std::isfinite(m_box[4 * i + 1]) || !std::isfinite(m_box[4 * i + 2]) || !std::isfinite(m_box[4 * i + 3]);
    }
    if (odd) {  // no culling at all for a tile holding a degenerate or non-finite anchor
      bx0 = by0 = -inf; bx1 = by1 = inf; amin = 0.0f; amax = inf;
    }
    t_bbox[4 * t + 0] = bx0; t_bbox[4 * t + 1] = by0; t_bbox[4 * t + 2] = bx1; t_bbox[4 * t + 3] = by1;
    t_area[2 * t + 0] = amin; t_area[2 * t + 1] = amax;
  }

  std::vector<ssdx_grid_combo> combos;
  std::vector<ssdx_grid_span> spans;
  std::vector<float> axes;
  std::vector<int> combo_span;
  detect_grids(corners_xyxy_host, area.data(), A, combos, spans, axes, &combo_span);
  const size_t nq = combos.size(), ns = spans.size(), nax = axes.size();

  ssdx_anchors* h = (ssdx_anchors*)calloc(1, sizeof(ssdx_anchors));
  ssdx_grid_combo* combos_host = nq ? (ssdx_grid_combo*)malloc(nq * sizeof(ssdx_grid_combo)) : nullptr;
  if (h == nullptr || (nq && combos_host == nullptr)) {
    free(h);
    free(combos_host);
    ssdx_set_error("ssdx_anchors_create: out of host memory");
    return SSDX_ERR_INVALID;
  }
  h->device = device;
  h->A = num_anchors;
  h->n_tiles = (int32_t)n_tiles;
  h->n_combos = (int32_t)nq;
  h->n_spans = (int32_t)ns;
  h->combos_host = combos_host;
  for (size_t i = 0; i < nq; ++i) combos_host[i] = combos[i];
  h->n_axes = (int32_t)nax;
  for (size_t i = 0; i < ns; ++i) {
    h->max_span_anchors = std::max(h->max_span_anchors, (int32_t)spans[i].n);
    h->max_span_combos = std::max(h->max_span_combos, (int32_t)spans[i].nb);
    const ssdx_grid_combo& first = combos[spans[i].combo0];
    const ssdx_grid_combo& last = combos[spans[i].combo0 + spans[i].nb - 1];
    const int map_axes = (last.ax + (2 * last.W + 2 * last.H + 3) / 4 * 4) - first.ax;
    h->max_map_axes = std::max(h->max_map_axes, (int32_t)map_axes);
  }
  cudaError_t e = cudaSuccess;
  // one allocation: corners | enc_cent | dec_cent | m_box | t_bbox | t_area | area | m_area | m_idx | combos | spans
  char* base = nullptr;
  const size_t bytes = A * (3 * sizeof(float4) + sizeof(float)) + Ap * (sizeof(float4) + sizeof(float) + sizeof(int)) +
                       n_tiles * (sizeof(float4) + sizeof(float2)) + 64 + nq * sizeof(ssdx_grid_combo) +
                       ns * sizeof(ssdx_grid_span) + nax * sizeof(float) + nq * sizeof(int2);
  e = cudaMalloc((void**)&base, bytes);
  if (e == cudaSuccess) {
    h->corners = (float4*)base;
    h->enc_cent = h->corners + A;
    h->dec_cent = h->enc_cent + A;
    h->m_box = h->dec_cent + A;
    h->t_bbox = h->m_box + Ap;
    h->t_area = (float2*)(h->t_bbox + n_tiles);
    h->area = (float*)(h->t_area + n_tiles);
    h->m_area = h->area + A;
    h->m_idx = (int*)(h->m_area + Ap);
    h->combos = (ssdx_grid_combo*)ssdx_align_up((size_t)(h->m_idx + Ap), 32);
    h->spans = (ssdx_grid_span*)(h->combos + nq);
    h->axes = (float*)(h->spans + ns);
    h->combo_span = (int2*)(h->axes + nax);  // nax is a multiple of 4 floats: 16-byte aligned
    e = cudaMemcpy(h->corners, corners_xyxy_host, A * sizeof(float4), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMemcpy(h->enc_cent, enc.data(), A * sizeof(float4), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->dec_cent, dec.data(), A * sizeof(float4), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->area, area.data(), A * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->m_box, m_box.data(), Ap * sizeof(float4), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->m_area, m_area.data(), Ap * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->m_idx, m_idx.data(), Ap * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->t_bbox, t_bbox.data(), n_tiles * sizeof(float4), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->t_area, t_area.data(), n_tiles * sizeof(float2), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && nq)
    e = cudaMemcpy(h->combos, combos.data(), nq * sizeof(ssdx_grid_combo), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && ns)
    e = cudaMemcpy(h->spans, spans.data(), ns * sizeof(ssdx_grid_span), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && nax) e = cudaMemcpy(h->axes, axes.data(), nax * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && nq)
    e = cudaMemcpy(h->combo_span, combo_span.data(), nq * sizeof(int2), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    ssdx_set_error("ssdx_anchors_create: CUDA error: %s", cudaGetErrorString(e));
    if (base) cudaFree(base);
    free(combos_host);
    free(h);
    return SSDX_ERR_CUDA;
  }
  *out = h;
  return SSDX_OK;
}

int ssdx_anchors_destroy(ssdx_anchors* anchors) {
  if (anchors == nullptr) return SSDX_OK;
  {
    // at interpreter teardown the context may already be gone: errors are ignored on purpose
    ssdx_device_guard guard(anchors->device);
    if (guard.err == cudaSuccess && anchors->corners) cudaFree(anchors->corners);
    (void)cudaGetLastError();
  }
  free(anchors->combos_host);
  free(anchors);
  return SSDX_OK;
}

/* Host-only probe of the grid detection (no device needed): what ssdx_anchors_create would find for this table. */
int ssdx_grid_probe(const float* corners_xyxy_host, int32_t num_anchors, int32_t* num_grids, int32_t* num_spans,
                    int32_t* grids_out, int32_t capacity) {
  SSDX_CHECK_ARG(corners_xyxy_host != nullptr && num_anchors > 0, "ssdx_grid_probe: bad anchor table");
  const size_t A = (size_t)num_anchors;
  std::vector<float> area(A);
  for (size_t a = 0; a < A; ++a) {
    volatile float hh = corners_xyxy_host[4 * a + 3] - corners_xyxy_host[4 * a + 1];
    volatile float ww = corners_xyxy_host[4 * a + 2] - corners_xyxy_host[4 * a + 0];
    hh = hh + 1.0f;
    ww = ww + 1.0f;
    volatile float ar = hh * ww;
    area[a] = ar;
  }
  std::vector<ssdx_grid_combo> combos;
  std::vector<ssdx_grid_span> spans;
  std::vector<float> axes;
  detect_grids(corners_xyxy_host, area.data(), A, combos, spans, axes);
  if (num_grids) *num_grids = (int32_t)combos.size();
  if (num_spans) *num_spans = (int32_t)spans.size();
  if (grids_out)
    for (size_t i = 0; i < combos.size() && (int32_t)i < capacity; ++i) {
      grids_out[4 * i + 0] = combos[i].off;
      grids_out[4 * i + 1] = combos[i].H;
      grids_out[4 * i + 2] = combos[i].W;
      grids_out[4 * i + 3] = combos[i].nb;
    }
  return SSDX_OK;
}

/* Host-only: the window ssdx_encode's grid drivers would enumerate for every (GT box, grid) of this table at the IoU
 * bounds t[i]; out (n_boxes, num_grids, 4) = (first row, rows, first column, columns), rows = 0 for an empty window. */
int ssdx_grid_window_probe(const float* corners_xyxy_host, int32_t num_anchors, const float* gt_xyxy_host,
                           const float* t_host, int32_t n_boxes, int32_t* out) {
  SSDX_CHECK_ARG(corners_xyxy_host && gt_xyxy_host && t_host && out && num_anchors > 0 && n_boxes >= 0,
                 "ssdx_grid_window_probe: bad argument");
  const size_t A = (size_t)num_anchors;
  std::vector<float> area(A);
  for (size_t a = 0; a < A; ++a) {
    volatile float hh = corners_xyxy_host[4 * a + 3] - corners_xyxy_host[4 * a + 1];
    volatile float ww = corners_xyxy_host[4 * a + 2] - corners_xyxy_host[4 * a + 0];
    hh = hh + 1.0f;
    ww = ww + 1.0f;
    volatile float ar = hh * ww;
    area[a] = ar;
  }
  std::vector<ssdx_grid_combo> combos;
  std::vector<ssdx_grid_span> spans;
  std::vector<float> axes;
  detect_grids(corners_xyxy_host, area.data(), A, combos, spans, axes);
  for (int32_t i = 0; i < n_boxes; ++i)
    for (size_t q = 0; q < combos.size(); ++q) {
      float4 gb;
      gb.x = gt_xyxy_host[4 * i + 0]; gb.y = gt_xyxy_host[4 * i + 1];
      gb.z = gt_xyxy_host[4 * i + 2]; gb.w = gt_xyxy_host[4 * i + 3];
      ssdx_grid_rect r{0, 0, 0, 0};
      if (!ssdx_grid_window(combos[q], gb, t_host[i], 0, combos[q].H, r)) r = ssdx_grid_rect{0, 0, 0, 0};
      int32_t* o = out + ((size_t)i * combos.size() + q) * 4;
      o[0] = r.r0; o[1] = r.nr; o[2] = r.c0; o[3] = r.nc;
    }
  return SSDX_OK;
}

/* 1 when the anchor table was recognised as a union of regular grids (encode then uses closed-form windows). */
int ssdx_anchors_grid_info(const ssdx_anchors* anchors, int32_t* num_grids, int32_t* num_spans) {
  SSDX_CHECK_ARG(anchors != nullptr, "ssdx_anchors_grid_info: anchors is NULL");
  if (num_grids) *num_grids = anchors->n_combos;
  if (num_spans) *num_spans = anchors->n_spans;
  return SSDX_OK;
}

int ssdx_anchors_count(const ssdx_anchors* anchors, int32_t* num_anchors) {
  SSDX_CHECK_ARG(anchors != nullptr && num_anchors != nullptr, "ssdx_anchors_count: NULL argument");
  *num_anchors = anchors->A;
  return SSDX_OK;
}

int ssdx_anchors_device(const ssdx_anchors* anchors, int* device) {
  SSDX_CHECK_ARG(anchors != nullptr && device != nullptr, "ssdx_anchors_device: NULL argument");
  *device = anchors->device;
  return SSDX_OK;
}

// ---------------------------------------------------------------------------------------------
// GT ingest (host).  datacoder.py:194-199: strip, split on "\r\n", decode_csv [int32,f32,f32,f32,f32];
// evaluators.py:111-121: csv.reader (any newline).  Flip: datacoder.py:202-203.
// ---------------------------------------------------------------------------------------------
static bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

int ssdx_parse_gt_csv(const char* text, size_t text_len, int crlf_only, float flip_width, int32_t* labels_host,
                      float* boxes_xyxy_host, int32_t capacity, int32_t* count) {
  SSDX_CHECK_ARG(count != nullptr, "ssdx_parse_gt_csv: count is NULL");
  *count = 0;
  SSDX_CHECK_ARG(text != nullptr || text_len == 0, "ssdx_parse_gt_csv: text is NULL");
  SSDX_CHECK_ARG(capacity >= 0 && (capacity == 0 || (labels_host && boxes_xyxy_host)),
                 "ssdx_parse_gt_csv: output arrays missing");
  size_t b = 0, e = text_len;
  while (b < e && is_space(text[b])) ++b;  // tf.strings.strip
  while (e > b && is_space(text[e - 1])) --e;
  int32_t n = 0;
  size_t pos = b;
  char buf[64];
  while (pos < e) {
    // record end
    size_t r = pos;
    size_t next;
    if (crlf_only) {
      while (r < e && !(text[r] == '\r' && r + 1 < e && text[r + 1] == '\n')) ++r;
      next = (r < e) ? r + 2 : e;
    } else {
      while (r < e && text[r] != '\n' && text[r] != '\r') ++r;
      next = r;
      if (next < e && text[next] == '\r') ++next;
      if (next < e && text[next] == '\n') ++next;
    }
    if (r == pos && !crlf_only) {  // blank line (csv.reader would yield an empty row -> reference raises)
      ssdx_set_error("ssdx_parse_gt_csv: empty record at byte %zu", pos);
      return SSDX_ERR_INVALID;
    }
    // 5 comma-separated fields
    double v[5];
    size_t f = pos;
    for (int k = 0; k < 5; ++k) {
      size_t g = f;
      while (g < r && text[g] != ',') ++g;
      if ((k < 4 && g >= r) || (k == 4 && g != r)) {
        ssdx_set_error("ssdx_parse_gt_csv: record %d does not have 5 fields", n);
        return SSDX_ERR_INVALID;
      }
      size_t len = g - f;
      if (len == 0 || len >= sizeof(buf)) {
        ssdx_set_error("ssdx_parse_gt_csv: bad field %d in record %d", k, n);
        return SSDX_ERR_INVALID;
      }
      memcpy(buf, text + f, len);
      buf[len] = 0;
      char* endp = nullptr;
      v[k] = (k == 0) ? (double)strtol(buf, &endp, 10) : strtod(buf, &endp);
      while (endp && *endp && is_space(*endp)) ++endp;
      if (endp == buf || (endp && *endp)) {
        ssdx_set_error("ssdx_parse_gt_csv: cannot parse field %d of record %d ('%s')", k, n, buf);
        return SSDX_ERR_INVALID;
      }
      f = g + 1;
    }
    if (n < capacity) {
      labels_host[n] = (int32_t)v[0];
      volatile float x0 = (float)v[1], y0 = (float)v[2], x1 = (float)v[3], y1 = (float)v[4];
      if (flip_width > 0.0f) {
        volatile float nx0 = flip_width - x1, nx1 = flip_width - x0;
        x0 = nx0;
        x1 = nx1;
      }
      boxes_xyxy_host[4 * n + 0] = x0;
      boxes_xyxy_host[4 * n + 1] = y0;
      boxes_xyxy_host[4 * n + 2] = x1;
      boxes_xyxy_host[4 * n + 3] = y1;
    }
    ++n;
    pos = next;
  }
  *count = n;
  if (n > capacity) {
    ssdx_set_error("ssdx_parse_gt_csv: %d records but capacity is %d", n, capacity);
    return SSDX_ERR_WORKSPACE;
  }
  return SSDX_OK;
}


// ---------------------------------------------------------------------------------------------
// Batched GT ingest (host, multithreaded): the step right before encode / eval.
// One file per image (datacoder.py:194 tf.io.read_file, evaluators.py:111 open()), parsed by
// ssdx_parse_gt_csv into padded arrays labels (N, gmax), boxes (N, gmax, 4), counts (N).
// ---------------------------------------------------------------------------------------------
int ssdx_read_gt_csv_batch(const char* const* paths, int32_t n_files, int crlf_only, const uint8_t* flip,
                           float flip_width, int32_t gmax, int32_t* labels_host, float* boxes_xyxy_host,
                           int32_t* counts_host, int32_t threads) {
  SSDX_CHECK_ARG(n_files >= 0 && gmax >= 0, "ssdx_read_gt_csv_batch: n_files and gmax must be >= 0");
  SSDX_CHECK_ARG(n_files == 0 || (paths != nullptr && counts_host != nullptr),
                 "ssdx_read_gt_csv_batch: paths / counts_host is NULL");
  SSDX_CHECK_ARG(gmax == 0 || (labels_host != nullptr && boxes_xyxy_host != nullptr),
                 "ssdx_read_gt_csv_batch: output arrays missing");
  if (n_files == 0) return SSDX_OK;
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, (int)n_files));
  std::atomic<int32_t> next(0);
  std::atomic<int> first_error(0);  // 0 none, else status of the first failure
  std::string error_text;
  std::atomic_flag error_lock = ATOMIC_FLAG_INIT;
  std::atomic<int> overflow(0);
  auto work = [&]() {
    std::vector<char> buf;
    for (;;) {
      const int32_t i = next.fetch_add(1);
      if (i >= n_files || first_error.load() != 0) return;
      int32_t* lab = gmax ? labels_host + (size_t)i * gmax : nullptr;
      float* box = gmax ? boxes_xyxy_host + (size_t)i * gmax * 4 : nullptr;
      if (gmax) {
        memset(lab, 0, (size_t)gmax * sizeof(int32_t));
        memset(box, 0, (size_t)gmax * 4 * sizeof(float));
      }
      counts_host[i] = 0;
      int st = SSDX_OK;
      std::string msg;
      FILE* f = paths[i] ? fopen(paths[i], "rb") : nullptr;
      if (f == nullptr) {
        st = SSDX_ERR_INVALID;
        msg = std::string("ssdx_read_gt_csv_batch: cannot open '") + (paths[i] ? paths[i] : "(null)") + "'";
      } else {
        fseek(f, 0, SEEK_END);
        const long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        buf.resize(sz > 0 ? (size_t)sz : 0);
        const size_t got = sz > 0 ? fread(buf.data(), 1, (size_t)sz, f) : 0;
        fclose(f);
        int32_t n = 0;
        st = ssdx_parse_gt_csv(buf.data(), got, crlf_only, (flip && flip[i]) ? flip_width : 0.0f, lab, box, gmax, &n);
        counts_host[i] = n;
        if (st == SSDX_ERR_WORKSPACE) {  // more records than gmax: counts hold the true number, caller retries
          overflow.store(1);
          st = SSDX_OK;
        } else if (st != SSDX_OK) {
          msg = std::string(ssdx_last_error()) + " (file '" + paths[i] + "')";
        }
      }
      if (st != SSDX_OK) {
        while (error_lock.test_and_set()) {}
        if (first_error.load() == 0) {
          error_text = msg;
          first_error.store(st);
        }
        error_lock.clear();
        return;
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  if (first_error.load() != 0) {
    ssdx_set_error("%s", error_text.c_str());
    return first_error.load();
  }
  if (overflow.load()) {
    int32_t need = 0;
    for (int32_t i = 0; i < n_files; ++i) need = std::max(need, counts_host[i]);
    ssdx_set_error("ssdx_read_gt_csv_batch: a file holds %d records but gmax is %d", need, gmax);
    return SSDX_ERR_WORKSPACE;
  }
  return SSDX_OK;
}

}  // extern "C"
