// Data-parallel gradient exchange of the training step: NCCL all-reduce (average) of the flat fp32 gradient buffer in a few
// buckets, each launched on a dedicated communication stream as soon as the backward pass has produced its last gradient, so
// the exchange of the head / late blocks runs under the BatchNorm-backward / dgrad chain of the early (large) layers.
// The reference trains on one GPU (train/train.py:155-171, SURVEY.md §5 "Distributed communication backend: none"); this is the
// one collective BASELINE.json's north_star adds: "NCCL allreduce over NVLink of bucketed gradients overlapped with backward".
//
// NCCL is resolved at run time from the process (torch has already loaded its bundled libnccl.so.2; the system library is the
// fallback): no NCCL header or link-time dependency, the five entry points used are declared here with their public signatures.
// The communicator is this library's own (ncclCommInitRank with an id the host layer broadcasts), one per process = per GPU.
// Everything is enqueued (stream-ordered, capturable in a CUDA graph); nothing here synchronises the device.
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "net.h"

namespace mtgseg {

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } NcclUniqueId;  // NCCL_UNIQUE_ID_BYTES
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(ncclComm_t*, int, NcclUniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*CommDestroyFn)(ncclComm_t);
typedef const char* (*GetErrorStringFn)(int);
constexpr int kNcclFloat32 = 7, kNcclAvg = 4;  // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since 2.10)

struct Nccl {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn error_string = nullptr;
  bool ok = false;
};

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    n.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the process already uses (torch's)
    if (!n.handle) n.handle = dlopen("libnccl.so.2", RTLD_NOW);
    if (!n.handle) return;
    n.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(n.handle, "ncclGetUniqueId"));
    n.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(n.handle, "ncclCommInitRank"));
    n.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(n.handle, "ncclAllReduce"));
    n.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(n.handle, "ncclCommDestroy"));
    n.error_string = reinterpret_cast<GetErrorStringFn>(dlsym(n.handle, "ncclGetErrorString"));
    n.ok = n.get_unique_id && n.comm_init_rank && n.all_reduce && n.comm_destroy;
  });
  return n;
}

struct DpState {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = -1;
  cudaStream_t stream = nullptr;  // communication stream
  cudaEvent_t ev_main = nullptr, ev_side = nullptr, ev_done = nullptr;
};
DpState g_dp;
std::mutex g_dp_mu;

#define MTG_NCCL(expr)                                                                                       \
  do {                                                                                                       \
    const int _r = (expr);                                                                                   \
    if (_r != 0) {                                                                                           \
      set_error("%s failed: %s", #expr, nccl().error_string ? nccl().error_string(_r) : "NCCL error");       \
      return MTG_ERR_CUDA;                                                                                   \
    }                                                                                                        \
  } while (0)

}  // namespace

int dp_unique_id(void* out128) {
  MTG_REQUIRE(out128 != nullptr, MTG_ERR_ARG, "dp_unique_id: null pointer");
  MTG_REQUIRE(nccl().ok, MTG_ERR_UNSUPPORTED, "libnccl.so.2 could not be resolved in this process");
  NcclUniqueId id;
  MTG_NCCL(nccl().get_unique_id(&id));
  memcpy(out128, &id, sizeof(id));
  return MTG_OK;
}

int dp_init(const void* id128, int rank, int world) {
  MTG_REQUIRE(id128 != nullptr && world >= 1 && rank >= 0 && rank < world, MTG_ERR_ARG, "dp_init: bad arguments");
  MTG_REQUIRE(nccl().ok, MTG_ERR_UNSUPPORTED, "libnccl.so.2 could not be resolved in this process");
  std::lock_guard<std::mutex> lock(g_dp_mu);
  MTG_REQUIRE(g_dp.comm == nullptr, MTG_ERR_ARG, "dp_init: already initialised (call mtgseg_dp_shutdown first)");
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  MTG_CUDA(cudaGetDevice(&g_dp.device));
  MTG_NCCL(nccl().comm_init_rank(&g_dp.comm, world, id, rank));
  MTG_CUDA(cudaStreamCreateWithFlags(&g_dp.stream, cudaStreamNonBlocking));
  MTG_CUDA(cudaEventCreateWithFlags(&g_dp.ev_main, cudaEventDisableTiming));
  MTG_CUDA(cudaEventCreateWithFlags(&g_dp.ev_side, cudaEventDisableTiming));
  MTG_CUDA(cudaEventCreateWithFlags(&g_dp.ev_done, cudaEventDisableTiming));
  g_dp.rank = rank;
  g_dp.world = world;
  return MTG_OK;
}

int dp_world() { return g_dp.comm ? g_dp.world : 0; }

int dp_shutdown() {
  std::lock_guard<std::mutex> lock(g_dp_mu);
  if (!g_dp.comm) return MTG_OK;
  cudaStreamSynchronize(g_dp.stream);
  nccl().comm_destroy(g_dp.comm);
  cudaStreamDestroy(g_dp.stream);
  cudaEventDestroy(g_dp.ev_main); cudaEventDestroy(g_dp.ev_side); cudaEventDestroy(g_dp.ev_done);
  g_dp = DpState{};
  return MTG_OK;
}

// One bucket: when everything enqueued so far on `main` (and on `side`, if given) has finished, average buf[0..n) over the ranks
// on the communication stream.  dp_join makes `main` wait for all buckets fired so far.
int dp_fire_bucket(float* buf, size_t n, cudaStream_t main, cudaStream_t side) {
  MTG_REQUIRE(g_dp.comm != nullptr, MTG_ERR_ARG, "data-parallel exchange requested but mtgseg_dp_init was not called");
  if (n == 0) return MTG_OK;
  MTG_CUDA(cudaEventRecord(g_dp.ev_main, main));
  MTG_CUDA(cudaStreamWaitEvent(g_dp.stream, g_dp.ev_main, 0));
  if (side) {
    MTG_CUDA(cudaEventRecord(g_dp.ev_side, side));
    MTG_CUDA(cudaStreamWaitEvent(g_dp.stream, g_dp.ev_side, 0));
  }
  MTG_NCCL(nccl().all_reduce(buf, buf, n, kNcclFloat32, kNcclAvg, g_dp.comm, g_dp.stream));
  return MTG_OK;
}

int dp_join(cudaStream_t main) {
  MTG_REQUIRE(g_dp.comm != nullptr, MTG_ERR_ARG, "data-parallel exchange requested but mtgseg_dp_init was not called");
  MTG_CUDA(cudaEventRecord(g_dp.ev_done, g_dp.stream));
  MTG_CUDA(cudaStreamWaitEvent(main, g_dp.ev_done, 0));
  return MTG_OK;
}

}  // namespace mtgseg
