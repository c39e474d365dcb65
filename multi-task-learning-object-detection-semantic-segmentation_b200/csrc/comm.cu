// The only collective of the path (BASELINE config 5): the per-rank (C, 3) int64 table of
// (TP, FP, n_gt) counts that ssdx_eval_tp produces (evaluators.py:106-157 per shard) is all-gathered over the ranks of
// one box through NCCL (NVLink / NVSwitch), so that every rank can form the global counts of evaluators.py:159-187.
//
// libssdx does not link against NCCL: the process that evaluates on several GPUs already has one loaded
// (torch.distributed's libnccl.so.2), and the entry points below bind to it at first use with dlopen/dlsym
// (RTLD_NOLOAD first; then $SSDX_NCCL_LIB, libnccl.so.2, libnccl.so).  Without NCCL they fail with SSDX_ERR_NCCL; no
// other entry point depends on it.  The few NCCL declarations used are restated here (nccl.h 2.x ABI).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];  // NCCL_UNIQUE_ID_BYTES
} ncclUniqueId;
typedef int ncclResult_t;            // 0 = ncclSuccess
constexpr int kNcclInt64 = 4;        // ncclDataType_t::ncclInt64
constexpr int kNcclUint8 = 1, kNcclInt32 = 2, kNcclFloat32 = 7;

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
  const char* names[4] = {getenv("SSDX_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the process already uses, if any
  for (int i = 0; h == nullptr && i < 3; ++i)
    if (names[i] != nullptr && names[i][0] != 0) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) return;
  NcclApi a;
  a.lib = h;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
  a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
  a.GroupStart = (decltype(a.GroupStart))dlsym(h, "ncclGroupStart");
  a.GroupEnd = (decltype(a.GroupEnd))dlsym(h, "ncclGroupEnd");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
  a.GetVersion = (decltype(a.GetVersion))dlsym(h, "ncclGetVersion");
  if (a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.GroupStart && a.GroupEnd) g_nccl = a;
}

const NcclApi* nccl() {
  std::call_once(g_nccl_once, load_nccl);
  return g_nccl.lib != nullptr ? &g_nccl : nullptr;
}

int nccl_fail(const char* what, ncclResult_t r) {
  const NcclApi* a = nccl();
  ssdx_set_error("%s failed: %s (NCCL result %d)", what,
                 (a && a->GetErrorString) ? a->GetErrorString(r) : "unknown error", (int)r);
  return SSDX_ERR_NCCL;
}

__global__ void ssdx_sum_counts_kernel(const long long* __restrict__ gathered, const int nranks, const int n,
                                       long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long s = 0;
  for (int r = 0; r < nranks; ++r) s += gathered[(size_t)r * n + i];
  out[i] = s;
}

}  // namespace

struct ssdx_comm {
  ncclComm_t comm;
  int nranks, rank, device;
};

#define SSDX_NCCL_REQUIRED(a)                                                                              \
  do {                                                                                                     \
    if ((a) == nullptr) {                                                                                  \
      ssdx_set_error("NCCL is not available in this process (libnccl.so.2 not loaded; set SSDX_NCCL_LIB)"); \
      return SSDX_ERR_NCCL;                                                                                \
    }                                                                                                      \
  } while (0)

extern "C" {

int ssdx_nccl_version(int* version) {
  SSDX_CHECK_ARG(version != nullptr, "ssdx_nccl_version: version is NULL");
  *version = 0;
  const NcclApi* a = nccl();
  SSDX_NCCL_REQUIRED(a);
  if (a->GetVersion) {
    const ncclResult_t r = a->GetVersion(version);
    if (r != 0) return nccl_fail("ncclGetVersion", r);
  }
  return SSDX_OK;
}

int ssdx_comm_unique_id(uint8_t* id128) {
  SSDX_CHECK_ARG(id128 != nullptr, "ssdx_comm_unique_id: id128 is NULL");
  const NcclApi* a = nccl();
  SSDX_NCCL_REQUIRED(a);
  ncclUniqueId id;
  const ncclResult_t r = a->GetUniqueId(&id);
  if (r != 0) return nccl_fail("ncclGetUniqueId", r);
  memcpy(id128, id.internal, sizeof(id.internal));
  return SSDX_OK;
}

int ssdx_comm_create(const uint8_t* id128, int32_t nranks, int32_t rank, int device, ssdx_comm** out) {
  SSDX_CHECK_ARG(out != nullptr, "ssdx_comm_create: out is NULL");
  *out = nullptr;
  SSDX_CHECK_ARG(id128 != nullptr && nranks >= 1 && rank >= 0 && rank < nranks,
                 "ssdx_comm_create: bad id / rank %d of %d", rank, nranks);
  const NcclApi* a = nccl();
  SSDX_NCCL_REQUIRED(a);
  SSDX_DEVICE(device);
  ncclUniqueId id;
  memcpy(id.internal, id128, sizeof(id.internal));
  ssdx_comm* c = (ssdx_comm*)calloc(1, sizeof(ssdx_comm));
  SSDX_CHECK_ARG(c != nullptr, "ssdx_comm_create: out of host memory");
  const ncclResult_t r = a->CommInitRank(&c->comm, nranks, id, rank);  // collective: every rank calls it
  if (r != 0) {
    free(c);
    return nccl_fail("ncclCommInitRank", r);
  }
  c->nranks = nranks;
  c->rank = rank;
  c->device = device;
  *out = c;
  return SSDX_OK;
}

int ssdx_comm_destroy(ssdx_comm* comm) {
  if (comm == nullptr) return SSDX_OK;
  const NcclApi* a = nccl();
  if (a != nullptr && comm->comm != nullptr) {
    ssdx_device_guard guard(comm->device);
    if (guard.err == cudaSuccess) a->CommDestroy(comm->comm);
  }
  free(comm);
  return SSDX_OK;
}

int ssdx_allgather_counts(ssdx_comm* comm, const int64_t* counts, int64_t* gathered, int32_t num_classes,
                          ssdx_stream stream) {
  SSDX_CHECK_ARG(comm != nullptr && counts != nullptr && gathered != nullptr && num_classes >= 1,
                 "ssdx_allgather_counts: bad argument");
  const NcclApi* a = nccl();
  SSDX_NCCL_REQUIRED(a);
  SSDX_DEVICE(comm->device);
  const ncclResult_t r = a->AllGather(counts, gathered, (size_t)num_classes * 3, kNcclInt64, comm->comm,
                                      (cudaStream_t)stream);
  if (r != 0) return nccl_fail("ncclAllGather", r);
  return SSDX_OK;
}

int ssdx_allgather_eval(ssdx_comm* comm, const int64_t* counts, int64_t* gathered_counts, int32_t num_classes,
                        const uint8_t* tp, const float* confidence, const int32_t* label, int64_t n_local,
                        uint8_t* all_tp, float* all_confidence, int32_t* all_label, ssdx_stream stream) {
  SSDX_CHECK_ARG(comm != nullptr && counts != nullptr && gathered_counts != nullptr && num_classes >= 1,
                 "ssdx_allgather_eval: bad argument");
  SSDX_CHECK_ARG(n_local >= 0 && (n_local == 0 || (tp && confidence && label && all_tp && all_confidence && all_label)),
                 "ssdx_allgather_eval: prediction arrays are NULL");
  const NcclApi* a = nccl();
  SSDX_NCCL_REQUIRED(a);
  SSDX_DEVICE(comm->device);
  cudaStream_t s = (cudaStream_t)stream;
  ncclResult_t r = a->GroupStart();  // the four gathers leave as one NCCL launch
  if (r != 0) return nccl_fail("ncclGroupStart", r);
  ncclResult_t first = a->AllGather(counts, gathered_counts, (size_t)num_classes * 3, kNcclInt64, comm->comm, s);
  if (n_local > 0) {
    if (first == 0) first = a->AllGather(tp, all_tp, (size_t)n_local, kNcclUint8, comm->comm, s);
    if (first == 0) first = a->AllGather(confidence, all_confidence, (size_t)n_local, kNcclFloat32, comm->comm, s);
    if (first == 0) first = a->AllGather(label, all_label, (size_t)n_local, kNcclInt32, comm->comm, s);
  }
  r = a->GroupEnd();
  if (first != 0) return nccl_fail("ncclAllGather", first);
  if (r != 0) return nccl_fail("ncclGroupEnd", r);
  return SSDX_OK;
}

int ssdx_sum_counts(const int64_t* gathered, int32_t nranks, int32_t num_classes, int64_t* out, int device,
                    ssdx_stream stream) {
  SSDX_CHECK_ARG(gathered != nullptr && out != nullptr && nranks >= 1 && num_classes >= 1,
                 "ssdx_sum_counts: bad argument");
  SSDX_DEVICE(device);
  const int n = num_classes * 3;
  ssdx_sum_counts_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const long long*)gathered, nranks, n,
                                                                           (long long*)out);
  SSDX_LAUNCH_CHECK("ssdx_sum_counts_kernel");
  return SSDX_OK;
}

}  // extern "C"
