// NCCL helper behind the C ABI (comm_init_from_unique_id / allreduce_flat): the one exchange
// of the data-parallel path is a SUM all-reduce of the flat fp32 gradient buffer (2.35 MB) over
// NVLink 5 / NVSwitch, plus the start-up broadcast of rank 0's parameters.  Replaces what
// DistributedDataParallel does for the reference (src/mimic_runner.py:141-143, src/utils/main_util.py:43-62).
// NCCL is resolved at run time (dlopen of the libnccl.so.2 that torch already loaded, or the system one):
// the library has no link-time dependency on it and single-GPU use never touches it.
#include <dlfcn.h>

#include "common.cuh"

namespace ghnd {

struct NcclUniqueId {
  char internal[128];
};
typedef struct ncclComm* NcclComm;
enum { kNcclFloat32 = 7, kNcclSum = 0, kNcclSuccess = 0 };

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*);
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int);
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
  int (*CommDestroy)(NcclComm);
  const char* (*GetErrorString)(int);
  bool ok;
};

static NcclApi* nccl_api() {
  static NcclApi api = [] {
    NcclApi a;
    memset(&a, 0, sizeof(a));
    void* h = nullptr;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) return a;
    a.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(h, "ncclCommInitRank");
    a.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllReduce");
    a.Broadcast = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(h, "ncclBroadcast");
    a.CommDestroy = (int (*)(NcclComm))dlsym(h, "ncclCommDestroy");
    a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.Broadcast && a.CommDestroy;
    return a;
  }();
  return &api;
}

static int nccl_fail(int rc, const char* what) {
  NcclApi* a = nccl_api();
  set_error("%s: NCCL error %d (%s)", what, rc, a->GetErrorString ? a->GetErrorString(rc) : "?");
  return GHND_ERR_CUDA;
}

}  // namespace ghnd

struct ghnd_comm {
  ghnd::NcclComm comm;
  int world, rank;
};

extern "C" {
using namespace ghnd;

int ghnd_comm_unique_id(void* id128) {
  GHND_CHECK_ARG(id128 != nullptr, "comm_unique_id: null argument");
  NcclApi* a = nccl_api();
  GHND_CHECK_ARG(a->ok, "comm: libnccl.so.2 could not be loaded");
  const int rc = a->GetUniqueId(reinterpret_cast<NcclUniqueId*>(id128));
  return rc == kNcclSuccess ? GHND_OK : nccl_fail(rc, "ncclGetUniqueId");
}

int ghnd_comm_init_from_unique_id(const void* id128, int world_size, int rank, ghnd_comm_t** out) {
  GHND_CHECK_ARG(id128 && out, "comm_init_from_unique_id: null argument");
  GHND_CHECK_ARG(world_size >= 1 && rank >= 0 && rank < world_size, "comm_init: bad rank %d of %d", rank, world_size);
  *out = nullptr;
  NcclApi* a = nccl_api();
  GHND_CHECK_ARG(a->ok, "comm: libnccl.so.2 could not be loaded");
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NcclComm c = nullptr;
  const int rc = a->CommInitRank(&c, world_size, id, rank);  // binds the calling thread's current device
  if (rc != kNcclSuccess) return nccl_fail(rc, "ncclCommInitRank");
  ghnd_comm* h = new ghnd_comm();
  h->comm = c;
  h->world = world_size;
  h->rank = rank;
  *out = h;
  return GHND_OK;
}

int ghnd_comm_allreduce_flat(ghnd_comm_t* comm, float* buf, int64_t n, void* stream) {
  GHND_CHECK_ARG(comm && buf && n > 0, "comm_allreduce_flat: bad argument");
  const int rc = nccl_api()->AllReduce(buf, buf, (size_t)n, kNcclFloat32, kNcclSum, comm->comm, (cudaStream_t)stream);
  return rc == kNcclSuccess ? GHND_OK : nccl_fail(rc, "ncclAllReduce");
}

int ghnd_comm_broadcast_flat(ghnd_comm_t* comm, float* buf, int64_t n, int root, void* stream) {
  GHND_CHECK_ARG(comm && buf && n > 0 && root >= 0 && root < comm->world, "comm_broadcast_flat: bad argument");
  const int rc = nccl_api()->Broadcast(buf, buf, (size_t)n, kNcclFloat32, root, comm->comm, (cudaStream_t)stream);
  return rc == kNcclSuccess ? GHND_OK : nccl_fail(rc, "ncclBroadcast");
}

void ghnd_comm_destroy(ghnd_comm_t* comm) {
  if (comm == nullptr) return;
  if (comm->comm) nccl_api()->CommDestroy(comm->comm);
  delete comm;
}
}
