// comm.cu -- the library's multi-GPU plumbing: one process per GPU, one NCCL communicator per handle.
//
// The PSULVSB path exchanges very little: the 8-byte best-hypothesis key of a sharded scoring sweep
// (ncclAllReduce max), per-row popcounts / per-rank edge counts of a row-sharded consistency stage, and the compacted
// edge lists themselves when one large registration is solved by several GPUs (in-place all-gather-v as grouped
// broadcasts).  All calls are issued in-stream; nothing here synchronises the host.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, reusing a copy the process already holds -- e.g. PyTorch's), so
// single-GPU users of libpsulvsb_b200.so carry no NCCL dependency.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstring>
#include <mutex>
#include <string>

#include "common.cuh"
#include "engine.cuh"

namespace psulvsb {

namespace {

struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  std::string error;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // a copy already in the process
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      const char* e = dlerror();
      api.error = std::string("libnccl.so.2 not found: ") + (e ? e : "");
      return;
    }
    api.lib = h;
    bool ok = true;
    auto sym = [&](const char* name) -> void* {
      void* p = dlsym(h, name);
      if (!p) {
        ok = false;
        api.error = std::string("libnccl lacks ") + name;
      }
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) api.lib = nullptr;
  });
  return api;
}

int need_nccl() {
  if (!nccl().lib) return fail(PSULVSB_ERR_UNSUPPORTED, "multi-GPU entry point without NCCL: " + nccl().error);
  return PSULVSB_OK;
}

#define PSU_NCCL(call)                                                                                   \
  do {                                                                                                   \
    ncclResult_t r__ = (call);                                                                           \
    if (r__ != ncclSuccess)                                                                              \
      return ::psulvsb::fail(PSULVSB_ERR_CUDA, std::string(#call) + ": " + nccl().GetErrorString(r__)); \
  } while (0)

}  // namespace

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
};

static_assert(sizeof(ncclUniqueId) == PSULVSB_UNIQUE_ID_BYTES, "psulvsb.h states the size of ncclUniqueId");

int comm_unique_id(void* out128) {
  if (int rc = need_nccl()) return rc;
  ncclUniqueId id;
  PSU_NCCL(nccl().GetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
  return PSULVSB_OK;
}

int comm_create(Comm** out, int device, int rank, int world, const void* id128) {
  *out = nullptr;
  if (int rc = need_nccl()) return rc;
  if (world < 1 || rank < 0 || rank >= world || !id128) return fail(PSULVSB_ERR_INVALID, "comm_create: bad rank / world / id");
  PSU_CUDA(cudaSetDevice(device));
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  Comm* c = new Comm();
  c->rank = rank;
  c->world = world;
  c->device = device;
  ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return fail(PSULVSB_ERR_CUDA, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r));
  }
  *out = c;
  return PSULVSB_OK;
}

void comm_destroy(Comm* c) {
  if (!c) return;
  if (c->comm && nccl().lib) nccl().CommDestroy(c->comm);
  delete c;
}

int comm_rank(const Comm* c) { return c ? c->rank : 0; }
int comm_world(const Comm* c) { return c ? c->world : 1; }

int comm_allreduce_sum_u32(Comm* c, cudaStream_t st, uint32_t* d_inout, size_t n) {
  if (!c || c->world == 1) return PSULVSB_OK;
  PSU_NCCL(nccl().AllReduce(d_inout, d_inout, n, ncclUint32, ncclSum, c->comm, st));
  return PSULVSB_OK;
}

int comm_allreduce_max_u64(Comm* c, cudaStream_t st, unsigned long long* d_inout, size_t n) {
  if (!c || c->world == 1) return PSULVSB_OK;
  PSU_NCCL(nccl().AllReduce(d_inout, d_inout, n, ncclUint64, ncclMax, c->comm, st));
  return PSULVSB_OK;
}

int comm_allgather_u64(Comm* c, cudaStream_t st, const unsigned long long* d_send, unsigned long long* d_recv) {
  if (!c || c->world == 1) {
    PSU_CUDA(cudaMemcpyAsync(d_recv, d_send, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    return PSULVSB_OK;
  }
  PSU_NCCL(nccl().AllGather(d_send, d_recv, 1, ncclUint64, c->comm, st));
  return PSULVSB_OK;
}

// in-place all-gather-v: rank r owns elements [offsets[r], offsets[r + 1]) of `base` on every rank
int comm_allgatherv_inplace_u32(Comm* c, cudaStream_t st, uint32_t* base, const unsigned long long* offsets) {
  if (!c || c->world == 1) return PSULVSB_OK;
  PSU_NCCL(nccl().GroupStart());
  for (int r = 0; r < c->world; ++r) {
    const unsigned long long cnt = offsets[r + 1] - offsets[r];
    if (cnt == 0) continue;
    ncclResult_t e = nccl().Broadcast(base + offsets[r], base + offsets[r], (size_t)cnt, ncclUint32, r, c->comm, st);
    if (e != ncclSuccess) {
      nccl().GroupEnd();
      return fail(PSULVSB_ERR_CUDA, std::string("ncclBroadcast: ") + nccl().GetErrorString(e));
    }
  }
  PSU_NCCL(nccl().GroupEnd());
  return PSULVSB_OK;
}

// Row block [begin, end) of the upper-triangular pair set such that every rank owns about the same number of pairs
// (row i has n - 1 - i of them): boundaries at n (1 - sqrt(1 - k / world)).
void triangular_row_range(int n, int rank, int world, int* begin, int* end) {
  auto bound = [&](int k) -> int {
    if (k <= 0) return 0;
    if (k >= world) return n;
    return (int)std::llround((double)n * (1.0 - std::sqrt(1.0 - (double)k / (double)world)));
  };
  *begin = bound(rank);
  *end = bound(rank + 1);
}

}  // namespace psulvsb
