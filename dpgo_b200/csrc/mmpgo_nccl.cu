// NCCL transport of the sharded driver, called from C++ on the handle's stream: the boundary poses of
// inter-node loop closures (DPGOHash::receive wire format, C++/DPGO/src/DPGOHash.cpp:45-82) as grouped
// ncclSend / ncclRecv between the ranks that share an edge, and the scalar all-reduce of AMM-PGO*
// (DPGOStar.cpp:147-171).  libnccl is resolved at run time (dlopen: the copy the process already loaded, e.g.
// torch's, else the system one), so libmmpgo.so carries no link dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "mmpgo_driver.cuh"

namespace mmpgo {

namespace {
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_api;
std::once_flag g_once;

void load_api() {
  for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
    g_api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (g_api.lib) break;
  }
  if (!g_api.lib) return;
#define SYM(f) g_api.f = reinterpret_cast<decltype(g_api.f)>(dlsym(g_api.lib, "nccl" #f)); if (!g_api.f) return
  SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(CommAbort); SYM(GroupStart); SYM(GroupEnd); SYM(Send); SYM(Recv);
  SYM(AllReduce); SYM(GetErrorString);
#undef SYM
  g_api.ok = true;
}
bool api() {
  std::call_once(g_once, load_api);
  if (!g_api.ok) set_error("libnccl.so.2 not found or incomplete");
  return g_api.ok;
}
int fail(const char *what, ncclResult_t r) {
  set_error(std::string(what) + ": " + g_api.GetErrorString(r));
  return MMPGO_ERR_CUDA;
}
}  // namespace

int nccl_unique_id(void *id) {
  if (!api()) return MMPGO_ERR_UNSUPPORTED;
  ncclUniqueId u;
  ncclResult_t r = g_api.GetUniqueId(&u);
  if (r != ncclSuccess) return fail("ncclGetUniqueId", r);
  std::memcpy(id, &u, sizeof(u));
  return 0;
}

int nccl_init(Handle *h, const void *id) {
  if (!api()) return MMPGO_ERR_UNSUPPORTED;
  if ((int)h->send_poses.size() != h->world) { set_error("set_sharding first"); return MMPGO_ERR_STATE; }
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  ncclComm_t comm = nullptr;
  ncclResult_t r = g_api.CommInitRank(&comm, h->world, u, h->rank);
  if (r != ncclSuccess) return fail("ncclCommInitRank", r);
  h->nccl_comm = comm;
  return 0;
}

// ncclCommDestroy may wait for the peers' communicators; handles are destroyed whenever their owner lets go of
// them (a garbage collector on the Python side), not in lock step, so the local resources are released with
// ncclCommAbort once the handle's stream has drained.
void nccl_destroy(Handle *h) {
  if (h->nccl_comm && g_api.ok) {
    cudaStreamSynchronize(h->stream);
    g_api.CommAbort(static_cast<ncclComm_t>(h->nccl_comm));
  }
  h->nccl_comm = nullptr;
}

// per-peer chunks of `send` (send_counts doubles each, rank order) -> peers; their chunks -> `recv`
int nccl_exchange(Handle *h, const double *send, const int64_t *send_counts, double *recv, const int64_t *recv_counts) {
  ncclComm_t comm = static_cast<ncclComm_t>(h->nccl_comm);
  ncclResult_t r = g_api.GroupStart();
  if (r != ncclSuccess) return fail("ncclGroupStart", r);
  int64_t so = 0, ro = 0;
  for (int q = 0; q < h->world; ++q) {
    if (send_counts[q] > 0 && (r = g_api.Send(send + so, (size_t)send_counts[q], ncclFloat64, q, comm, h->stream)) != ncclSuccess) return fail("ncclSend", r);
    if (recv_counts[q] > 0 && (r = g_api.Recv(recv + ro, (size_t)recv_counts[q], ncclFloat64, q, comm, h->stream)) != ncclSuccess) return fail("ncclRecv", r);
    so += send_counts[q]; ro += recv_counts[q];
  }
  r = g_api.GroupEnd();
  if (r != ncclSuccess) return fail("ncclGroupEnd", r);
  return 0;
}

// in-place sum of n device doubles over all ranks, on the handle's stream
int nccl_allreduce(Handle *h, double *vals_dev, int n) {
  ncclResult_t r = g_api.AllReduce(vals_dev, vals_dev, (size_t)n, ncclFloat64, ncclSum, static_cast<ncclComm_t>(h->nccl_comm), h->stream);
  if (r != ncclSuccess) return fail("ncclAllReduce", r);
  return 0;
}

}  // namespace mmpgo
