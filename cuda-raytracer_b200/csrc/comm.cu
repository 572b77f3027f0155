// Multi-GPU combine of the per-GPU accumulation buffers behind the C ABI (b2rt_comm_*, b2rt_reduce_accum*).
//
// The reference is single-GPU (no cudaSetDevice / NCCL anywhere in src/cudaRenderer.cu); north_star asks for sample
// sharding with the scene replicated and ONE NCCL reduce of the per-GPU float4 (rgb sum, sample count) buffers.  This
// file is that reduce: ncclReduce(sum, fp32) enqueued on the renderer's own stream, so it is ordered behind the last
// k_accumulate of the frame without a host synchronisation.
//
// NCCL is bound at run time (dlopen / dlsym), not at link time: inside a process that already carries an NCCL (PyTorch
// bundles its own) the library must use THAT copy, and libb2rt.so must still load on a box without NCCL (every
// b2rt_comm_* call then fails with B2RT_ERR_INVALID "NCCL not available").
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "render.cuh"

struct b2rt_comm {
  ncclComm_t comm = nullptr;
  int device = 0, rank = 0, n_ranks = 1;
};

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // the copy the process already has (RTLD_NOLOAD), else the system's.  A host program that also uses a framework
    // with a bundled NCCL (PyTorch) must load the framework first: one process, one libnccl.so.2.
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (api.lib) break; }
    if (!api.lib)
      for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (api.lib) break; }
    if (!api.lib) return;
    auto sym = [&](const char* s) { return dlsym(api.lib, s); };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.Reduce = (decltype(api.Reduce))sym("ncclReduce");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.Reduce && api.AllReduce &&
             api.GroupStart && api.GroupEnd && api.GetErrorString;
  });
  return api;
}

int need_nccl() {
  if (nccl().ok) return B2RT_OK;
  b2rt::set_error("NCCL not available (libnccl.so.2 could not be loaded)");
  return B2RT_ERR_INVALID;
}

#define B2RT_NCCL_OK(call)                                                                            \
  do {                                                                                                \
    ncclResult_t r__ = (call);                                                                        \
    if (r__ != ncclSuccess) {                                                                         \
      b2rt::set_error(std::string(#call) + ": " + nccl().GetErrorString(r__));                        \
      return B2RT_ERR_CUDA;                                                                           \
    }                                                                                                 \
  } while (0)

static_assert(sizeof(ncclUniqueId) == B2RT_COMM_ID_BYTES, "b2rt.h B2RT_COMM_ID_BYTES must match ncclUniqueId");

}  // namespace

extern "C" {

int b2rt_comm_version(void) {
  int v = 0;
  if (!nccl().ok || !nccl().GetVersion || nccl().GetVersion(&v) != ncclSuccess) return 0;
  return v;
}

int b2rt_comm_unique_id(uint8_t id[B2RT_COMM_ID_BYTES]) {
  if (!id) { b2rt::set_error("null argument"); return B2RT_ERR_INVALID; }
  int rc = need_nccl();
  if (rc) return rc;
  ncclUniqueId u;
  B2RT_NCCL_OK(nccl().GetUniqueId(&u));
  memcpy(id, &u, sizeof u);
  return B2RT_OK;
}

int b2rt_comm_create(int32_t n_ranks, int32_t rank, const uint8_t id[B2RT_COMM_ID_BYTES], int32_t device, b2rt_comm** out) {
  if (!out || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) { b2rt::set_error("bad argument"); return B2RT_ERR_INVALID; }
  *out = nullptr;
  int rc = need_nccl();
  if (rc) return rc;
  if (device < 0) B2RT_CUDA_OK(cudaGetDevice(&device));
  B2RT_CUDA_OK(cudaSetDevice(device));
  b2rt_comm* c = new (std::nothrow) b2rt_comm();
  if (!c) return B2RT_ERR_OOM;
  c->device = device; c->rank = rank; c->n_ranks = n_ranks;
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  ncclResult_t r = nccl().CommInitRank(&c->comm, n_ranks, u, rank);
  if (r != ncclSuccess) { b2rt::set_error(std::string("ncclCommInitRank: ") + nccl().GetErrorString(r)); delete c; return B2RT_ERR_CUDA; }
  *out = c;
  return B2RT_OK;
}

int b2rt_comm_create_all(int32_t n, const int32_t* devices, b2rt_comm** out) {
  if (!out || n < 1 || n > 64) { b2rt::set_error("bad argument"); return B2RT_ERR_INVALID; }
  for (int i = 0; i < n; ++i) out[i] = nullptr;
  int rc = need_nccl();
  if (rc) return rc;
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) devs[i] = devices ? devices[i] : i;
  std::vector<ncclComm_t> comms(n, nullptr);
  B2RT_NCCL_OK(nccl().CommInitAll(comms.data(), n, devs.data()));
  for (int i = 0; i < n; ++i) {
    b2rt_comm* c = new (std::nothrow) b2rt_comm();
    if (!c) { for (int k = 0; k < n; ++k) { if (out[k]) { delete out[k]; out[k] = nullptr; } nccl().CommDestroy(comms[k]); } return B2RT_ERR_OOM; }
    c->comm = comms[i]; c->device = devs[i]; c->rank = i; c->n_ranks = n;
    out[i] = c;
  }
  return B2RT_OK;
}

void b2rt_comm_destroy(b2rt_comm* c) {
  if (!c) return;
  if (c->comm && nccl().ok) { cudaSetDevice(c->device); nccl().CommDestroy(c->comm); }
  delete c;
}

static int reduce_one(b2rt_renderer* h, b2rt_comm* c, int root) {
  b2rt::Renderer& R = h->r;
  if (!R.accum) { b2rt::set_error("no frame buffer"); return B2RT_ERR_INVALID; }
  if (R.device != c->device) { b2rt::set_error("renderer and communicator are on different devices"); return B2RT_ERR_INVALID; }
  B2RT_CUDA_OK(cudaSetDevice(R.device));
  const size_t n = (size_t)R.width * R.height * 4;
  // in place on the root; ordered behind the frame's last kernel on the renderer's stream
  if (root < 0) B2RT_NCCL_OK(nccl().AllReduce(R.accum, R.accum, n, ncclFloat, ncclSum, c->comm, R.stream));
  else B2RT_NCCL_OK(nccl().Reduce(R.accum, R.accum, n, ncclFloat, ncclSum, root, c->comm, R.stream));
  return B2RT_OK;
}

int b2rt_reduce_accum(b2rt_renderer* h, b2rt_comm* c, int32_t root) {
  if (!h || !c || root >= c->n_ranks) { b2rt::set_error("bad argument"); return B2RT_ERR_INVALID; }
  int rc = need_nccl();
  if (rc) return rc;
  if (h->r.running) { rc = h->r.wait(); if (rc) return rc; }   // overflowed waves are re-rendered before the sums leave the GPU
  return reduce_one(h, c, root);
}

int b2rt_reduce_accum_all(b2rt_renderer** hs, b2rt_comm** cs, int32_t n, int32_t root) {
  if (!hs || !cs || n < 1) { b2rt::set_error("bad argument"); return B2RT_ERR_INVALID; }
  int rc = need_nccl();
  if (rc) return rc;
  for (int i = 0; i < n; ++i) {
    if (!hs[i] || !cs[i] || root >= cs[i]->n_ranks) { b2rt::set_error("bad argument"); return B2RT_ERR_INVALID; }
    if (hs[i]->r.running) { rc = hs[i]->r.wait(); if (rc) return rc; }
  }
  B2RT_NCCL_OK(nccl().GroupStart());
  for (int i = 0; i < n; ++i) {
    rc = reduce_one(hs[i], cs[i], root);
    if (rc) { nccl().GroupEnd(); return rc; }
  }
  B2RT_NCCL_OK(nccl().GroupEnd());
  return B2RT_OK;
}

}  // extern "C"
