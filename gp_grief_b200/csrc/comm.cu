// Row-shard exchange behind the C ABI: the all-reduce of the packed statistics (A | r | s) and of the kernel-parameter gradient
//  for hosts that have no torch.distributed.  One communicator per process (rank = GPU); NCCL is loaded with dlopen
// at the first call, so libgrief_b200.so itself has no link-time dependency on it and single-GPU hosts never touch it.
// The exchange is ~134 MB per evaluation at p = 4096 (< 1 ms on NVLink 5 against >= 0.4 s of row work per rank): a plain
// ncclAllReduce on the caller's stream, no fused kernel.
#include <dlfcn.h>

#include <cstring>

#include "plan.h"

namespace grief {

namespace {
struct NcclUid { char internal[128]; };            // ncclUniqueId
typedef void* NcclCommT;
typedef int (*GetUniqueIdFn)(NcclUid*);
typedef int (*CommInitRankFn)(NcclCommT*, int, NcclUid, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, NcclCommT, cudaStream_t);
typedef int (*CommDestroyFn)(NcclCommT);
typedef const char* (*GetErrorStringFn)(int);
constexpr int kNcclFloat64 = 8, kNcclSum = 0;      // ncclDouble, ncclSum (nccl.h; stable since NCCL 2.0)

struct NcclApi {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn error_string = nullptr;
};

int load_nccl(const NcclApi** out) {
  static NcclApi api;                               // the library handle is process-wide by nature
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (api.handle) {
      api.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(api.handle, "ncclGetUniqueId"));
      api.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(api.handle, "ncclCommInitRank"));
      api.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(api.handle, "ncclAllReduce"));
      api.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(api.handle, "ncclCommDestroy"));
      api.error_string = reinterpret_cast<GetErrorStringFn>(dlsym(api.handle, "ncclGetErrorString"));
    }
  }
  if (!api.handle || !api.get_unique_id || !api.comm_init_rank || !api.all_reduce || !api.comm_destroy)
    return fail(GRIEF_ERR_NCCL, "NCCL (libnccl.so.2) could not be loaded: %s", api.handle ? "missing symbols" : dlerror());
  *out = &api;
  return GRIEF_OK;
}

int nccl_fail(const NcclApi* api, const char* what, int code) {
  return fail(GRIEF_ERR_NCCL, "%s failed: %s", what, api->error_string ? api->error_string(code) : "unknown NCCL error");
}
}  // namespace

struct Comm {
  const NcclApi* api = nullptr;
  NcclCommT comm = nullptr;
  int world = 0, rank = 0, device = 0;
};

int comm_unique_id(char* id_out) {
  const NcclApi* api = nullptr;
  int rc = load_nccl(&api);
  if (rc != GRIEF_OK) return rc;
  NcclUid uid;
  const int e = api->get_unique_id(&uid);
  if (e != 0) return nccl_fail(api, "ncclGetUniqueId", e);
  std::memcpy(id_out, uid.internal, sizeof(uid.internal));
  return GRIEF_OK;
}

int comm_create(Comm** out, const char* id, int world, int rank) {
  GRIEF_REQUIRE(world >= 1 && rank >= 0 && rank < world, "grief_comm_create: rank %d of %d", rank, world);
  const NcclApi* api = nullptr;
  int rc = load_nccl(&api);
  if (rc != GRIEF_OK) return rc;
  NcclUid uid;
  std::memcpy(uid.internal, id, sizeof(uid.internal));
  Comm* c = new Comm();
  c->api = api; c->world = world; c->rank = rank;
  cudaGetDevice(&c->device);
  const int e = api->comm_init_rank(&c->comm, world, uid, rank);
  if (e != 0) { delete c; return nccl_fail(api, "ncclCommInitRank", e); }
  *out = c;
  return GRIEF_OK;
}

int comm_allreduce_sum(Comm* c, double* buf, int64_t count, cudaStream_t stream) {
  GRIEF_REQUIRE(count >= 0 && (count == 0 || buf != nullptr), "grief_comm_allreduce_sum: count=%lld", (long long)count);
  int dev = -1;
  cudaGetDevice(&dev);
  GRIEF_REQUIRE(dev == c->device, "communicator was created on device %d, the current device is %d", c->device, dev);
  if (count == 0 || c->world == 1) return GRIEF_OK;
  const int e = c->api->all_reduce(buf, buf, (size_t)count, kNcclFloat64, kNcclSum, c->comm, stream);
  if (e != 0) return nccl_fail(c->api, "ncclAllReduce", e);
  return GRIEF_OK;
}

void comm_destroy(Comm* c) {
  if (!c) return;
  if (c->comm) c->api->comm_destroy(c->comm);
  delete c;
}

}  // namespace grief
