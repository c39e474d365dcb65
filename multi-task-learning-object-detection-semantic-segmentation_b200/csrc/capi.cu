// libssdx: library-level entry points, the anchor table and the host-side GT ingest.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

static thread_local char g_err[512] = "";

void ssdx_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool ssdx_debug_sync() {
  const char* e = getenv("SSDX_DEBUG_SYNC");
  return e != nullptr && e[0] == '1';
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

extern "C" {

int ssdx_version(void) { return SSDX_VERSION; }

const char* ssdx_last_error(void) { return g_err; }

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
  SSDX_CUDA(cudaSetDevice(device));
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
  ssdx_anchors* h = (ssdx_anchors*)calloc(1, sizeof(ssdx_anchors));
  SSDX_CHECK_ARG(h != nullptr, "ssdx_anchors_create: out of host memory");
  h->device = device;
  h->A = num_anchors;
  cudaError_t e = cudaSuccess;
  // one allocation: corners | enc_cent | dec_cent | area
  char* base = nullptr;
  const size_t bytes = A * (3 * sizeof(float4) + sizeof(float));
  e = cudaMalloc((void**)&base, bytes);
  if (e == cudaSuccess) {
    h->corners = (float4*)base;
    h->enc_cent = h->corners + A;
    h->dec_cent = h->enc_cent + A;
    h->area = (float*)(h->dec_cent + A);
    e = cudaMemcpy(h->corners, corners_xyxy_host, A * sizeof(float4), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMemcpy(h->enc_cent, enc.data(), A * sizeof(float4), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->dec_cent, dec.data(), A * sizeof(float4), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->area, area.data(), A * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    ssdx_set_error("ssdx_anchors_create: CUDA error: %s", cudaGetErrorString(e));
    if (base) cudaFree(base);
    free(h);
    return SSDX_ERR_CUDA;
  }
  *out = h;
  return SSDX_OK;
}

int ssdx_anchors_destroy(ssdx_anchors* anchors) {
  if (anchors == nullptr) return SSDX_OK;
  cudaSetDevice(anchors->device);
  if (anchors->corners) cudaFree(anchors->corners);
  free(anchors);
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

}  // extern "C"
