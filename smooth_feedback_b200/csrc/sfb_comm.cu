// sfb_comm.cu -- the ONE collective of the path: an all-gather of batch results over NCCL (NVLink 5 / NVSwitch), issued from
// inside the library so that a C / C++ host has a multi-GPU path too (SURVEY 8(b) sfb_allgather_results, 8(e)).
//
// Instances are independent, so every rank solves a contiguous shard and nothing but results is exchanged.  The result of a
// shard is a handful of arrays (x, y, obj, status, iter / delta, P); they are NOT packed: the arrays are sent as one grouped
// NCCL operation (ncclGroupStart ... ncclAllGather per array ... ncclGroupEnd), which NCCL fuses into a single launch.
// The exchange runs on the communicator's own stream, ordered after the work already enqueued on the handle's stream, so
// that the solve of the next step (double-buffered outputs) overlaps it; sfb_comm_wait orders the handle's stream after it.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2, preferring a copy the process already loaded, e.g. PyTorch's): the
// library has no link-time dependency on it and single-GPU users never touch it.
#include "sfb_internal.hpp"

#include <dlfcn.h>
#include <nccl.h>

using namespace sfbi;

namespace {

struct NcclApi
{
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi& nccl()
{
  static NcclApi api;
  if (api.lib || !api.error.empty()) return api;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { api.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return api; }
  auto sym = [&](const char* name) -> void* {
    void* p = dlsym(lib, name);
    if (!p && api.error.empty()) api.error = std::string("libnccl lacks ") + name;
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  if (api.error.empty()) api.lib = lib;
  return api;
}

}  // namespace

struct sfb_comm
{
  sfb_context* h = nullptr;
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
  cudaStream_t stream = nullptr;  // the exchange runs here
  cudaEvent_t ready = nullptr;    // results of the handle's stream are complete
  cudaEvent_t done = nullptr;     // the exchange is complete
};

#define SFB_NCCL(h, call)                                                                                           \
  do {                                                                                                              \
    ncclResult_t r__ = (call);                                                                                      \
    if (r__ != ncclSuccess) return fail(h, SFB_ERR_CUDA, "%s failed: %s", #call, nccl().GetErrorString(r__));       \
  } while (0)

extern "C" {

int sfb_comm_unique_id(void* out128)
{
  if (!out128) return SFB_ERR_INVALID_ARGUMENT;
  NcclApi& api = nccl();
  if (!api.lib) return fail(nullptr, SFB_ERR_CUDA, "%s", api.error.c_str());
  static_assert(sizeof(ncclUniqueId) == SFB_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  if (api.GetUniqueId(&id) != ncclSuccess) return fail(nullptr, SFB_ERR_CUDA, "ncclGetUniqueId failed");
  std::memcpy(out128, &id, sizeof(id));
  return SFB_OK;
}

int sfb_comm_create(sfb_handle_t h, int world, int rank, const void* id128, sfb_comm_t* out)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  if (!out || !id128 || world < 1 || rank < 0 || rank >= world) return fail(h, SFB_ERR_INVALID_ARGUMENT, "bad argument to sfb_comm_create");
  *out = nullptr;
  NcclApi& api = nccl();
  if (!api.lib) return fail(h, SFB_ERR_CUDA, "%s", api.error.c_str());
  SFB_CUDA(h, cudaSetDevice(h->device));
  auto* c = new sfb_comm();
  c->h = h; c->world = world; c->rank = rank;
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  ncclResult_t r = api.CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return fail(h, SFB_ERR_CUDA, "ncclCommInitRank failed: %s", api.GetErrorString(r));
  }
  bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    sfb_comm_destroy(c);
    return fail(h, SFB_ERR_CUDA, "communicator stream setup failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  *out = c;
  return SFB_OK;
}

int sfb_comm_destroy(sfb_comm_t c)
{
  if (!c) return SFB_OK;
  cudaSetDevice(c->h->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->comm) nccl().CommDestroy(c->comm);
  if (c->ready) cudaEventDestroy(c->ready);
  if (c->done) cudaEventDestroy(c->done);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return SFB_OK;
}

int sfb_allgather_results(sfb_comm_t c, int count, const void* const* send, void* const* recv, const size_t* bytes_per_rank)
{
  if (!c) return SFB_ERR_INVALID_ARGUMENT;
  sfb_context* h = c->h;
  if (count < 1 || !send || !recv || !bytes_per_rank) return fail(h, SFB_ERR_INVALID_ARGUMENT, "bad argument to sfb_allgather_results");
  for (int k = 0; k < count; ++k)
    if (!send[k] || !recv[k] || mem_space(send[k]) != 1 || mem_space(recv[k]) != 1)
      return fail(h, SFB_ERR_INVALID_ARGUMENT, "sfb_allgather_results takes device pointers (array %d)", k);
  NcclApi& api = nccl();
  SFB_CUDA(h, cudaSetDevice(h->device));
  // the exchange starts when everything already enqueued on the handle's stream (the solve) has finished
  SFB_CUDA(h, cudaEventRecord(c->ready, h->stream));
  SFB_CUDA(h, cudaStreamWaitEvent(c->stream, c->ready, 0));
  SFB_NCCL(h, api.GroupStart());
  for (int k = 0; k < count; ++k) {
    ncclResult_t r = api.AllGather(send[k], recv[k], bytes_per_rank[k], ncclChar, c->comm, c->stream);
    if (r != ncclSuccess) {
      api.GroupEnd();
      return fail(h, SFB_ERR_CUDA, "ncclAllGather failed: %s", api.GetErrorString(r));
    }
  }
  SFB_NCCL(h, api.GroupEnd());
  SFB_CUDA(h, cudaEventRecord(c->done, c->stream));
  return SFB_OK;
}

int sfb_comm_wait(sfb_comm_t c, int host_sync)
{
  if (!c) return SFB_ERR_INVALID_ARGUMENT;
  sfb_context* h = c->h;
  SFB_CUDA(h, cudaSetDevice(h->device));
  if (host_sync) SFB_CUDA(h, cudaEventSynchronize(c->done));
  else SFB_CUDA(h, cudaStreamWaitEvent(h->stream, c->done, 0));
  return SFB_OK;
}

}  // extern "C"
