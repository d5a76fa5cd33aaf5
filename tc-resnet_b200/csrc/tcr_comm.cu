// tcr_comm.cu — NCCL through dlopen: the gradient all-reduce of the data-parallel step.
// No reference counterpart (const.py:7 pins one device).  dlopen("libnccl.so.2") resolves to the library
// already mapped by torch when the host process imported it, otherwise to the system libnccl; the
// communicator is our own (ncclCommInitRank with an id the host distributes), so it is independent of
// torch.distributed's process group and can be enqueued on the caller's stream / captured in a graph.
#include "tcr_net.h"

#ifndef TCR_EMU
#include <dlfcn.h>
#endif
#include <stdio.h>
#include <string.h>

namespace tcr {

static char g_comm_err[256] = "";
const char* comm_error() { return g_comm_err; }

#ifdef TCR_EMU
int comm_unique_id(void*) { snprintf(g_comm_err, sizeof(g_comm_err), "no NCCL in the CPU emulator"); return TCR_ERR_COMM; }
int comm_init(tcr_handle*, const void*, int, int) { return comm_unique_id(nullptr); }
void comm_destroy(tcr_handle*) {}
int comm_allreduce_sum(tcr_handle*, float*, int64_t, cudaStream_t) { return TCR_ERR_COMM; }
int comm_p2p_export(tcr_handle*, void*) { return comm_unique_id(nullptr); }
int comm_p2p_attach(tcr_handle*, const void*, int, int) { return comm_unique_id(nullptr); }
void comm_p2p_destroy(tcr_handle*) {}
#else

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclSum = 0, kNcclFloat32 = 7 };

static struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

static int load_nccl() {
  if (g_nccl.lib) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) {
    snprintf(g_comm_err, sizeof(g_comm_err), "dlopen(libnccl.so.2) failed: %s", dlerror());
    return TCR_ERR_COMM;
  }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) {
    snprintf(g_comm_err, sizeof(g_comm_err), "libnccl is missing required symbols");
    g_nccl.lib = nullptr;
    return TCR_ERR_COMM;
  }
  return 0;
}

static int nccl_fail(const char* what, ncclResult_t r) {
  snprintf(g_comm_err, sizeof(g_comm_err), "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
  return TCR_ERR_COMM;
}

int comm_unique_id(void* id128) {
  if (load_nccl()) return TCR_ERR_COMM;
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != 0) return nccl_fail("ncclGetUniqueId", r);
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int comm_init(tcr_handle* h, const void* id128, int rank, int world) {
  if (load_nccl()) return TCR_ERR_COMM;
  comm_destroy(h);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  ncclResult_t r = g_nccl.CommInitRank(&c, world, id, rank);
  if (r != 0) return nccl_fail("ncclCommInitRank", r);
  h->comm = c;
  h->rank = rank;
  h->world = world;
  return 0;
}

void comm_p2p_destroy(tcr_handle* h);
void comm_destroy(tcr_handle* h) {
  comm_p2p_destroy(h);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)h->comm);
  h->comm = nullptr;
  h->world = 1;
  h->rank = 0;
}

// ---- peer-memory exchange: every rank maps every other rank's gradient buffer and flag array through CUDA IPC ----
// handles: [0, 64) cudaIpcMemHandle_t of `grads`, [64, 128) of `flags`.
static int cuda_fail(const char* what, cudaError_t e) {
  snprintf(g_comm_err, sizeof(g_comm_err), "%s: %s", what, cudaGetErrorString(e));
  return TCR_ERR_COMM;
}
int comm_p2p_export(tcr_handle* h, void* handles128) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  auto& p = h->p2p;
  cudaError_t e;
  if (!p.grads) {
    if ((e = cudaMalloc((void**)&p.grads, (size_t)2 * h->n_train * sizeof(float))) != cudaSuccess) return cuda_fail("cudaMalloc(p2p grads)", e);
    if ((e = cudaMalloc((void**)&p.flags, 64 * sizeof(unsigned))) != cudaSuccess) return cuda_fail("cudaMalloc(p2p flags)", e);
    cudaMemset(p.grads, 0, (size_t)2 * h->n_train * sizeof(float));
    cudaMemset(p.flags, 0, 64 * sizeof(unsigned));
    cudaDeviceSynchronize();
  }
  cudaIpcMemHandle_t hg, hf;
  if ((e = cudaIpcGetMemHandle(&hg, p.grads)) != cudaSuccess) return cuda_fail("cudaIpcGetMemHandle", e);
  if ((e = cudaIpcGetMemHandle(&hf, p.flags)) != cudaSuccess) return cuda_fail("cudaIpcGetMemHandle", e);
  memcpy(handles128, &hg, 64);
  memcpy((char*)handles128 + 64, &hf, 64);
  return 0;
}
int comm_p2p_attach(tcr_handle* h, const void* all_handles, int rank, int world) {
  auto& p = h->p2p;
  if (!p.grads || world < 1 || world > 8) { snprintf(g_comm_err, sizeof(g_comm_err), "p2p: export first, world <= 8"); return TCR_ERR_COMM; }
  for (int r = 0; r < world; ++r) {
    if (r == rank) { p.peer_grads[r] = p.grads; p.peer_flags[r] = p.flags; continue; }
    cudaIpcMemHandle_t hg, hf;
    memcpy(&hg, (const char*)all_handles + (size_t)r * 128, 64);
    memcpy(&hf, (const char*)all_handles + (size_t)r * 128 + 64, 64);
    cudaError_t e;
    if ((e = cudaIpcOpenMemHandle((void**)&p.peer_grads[r], hg, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess) return cuda_fail("cudaIpcOpenMemHandle(grads)", e);
    if ((e = cudaIpcOpenMemHandle((void**)&p.peer_flags[r], hf, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess) return cuda_fail("cudaIpcOpenMemHandle(flags)", e);
  }
  h->rank = rank;
  h->world = world;
  p.attached = 1;
  p.step = 0;
  return 0;
}
void comm_p2p_destroy(tcr_handle* h) {
  auto& p = h->p2p;
  for (int r = 0; r < 8; ++r) {
    if (p.peer_grads[r] && p.peer_grads[r] != p.grads) cudaIpcCloseMemHandle(p.peer_grads[r]);
    if (p.peer_flags[r] && p.peer_flags[r] != p.flags) cudaIpcCloseMemHandle(p.peer_flags[r]);
    p.peer_grads[r] = nullptr; p.peer_flags[r] = nullptr;
  }
  if (p.grads) cudaFree(p.grads);
  if (p.flags) cudaFree(p.flags);
  p.grads = nullptr; p.flags = nullptr; p.attached = 0;
}

int comm_allreduce_sum(tcr_handle* h, float* buf, int64_t count, cudaStream_t s) {
  ncclResult_t r = g_nccl.AllReduce(buf, buf, (size_t)count, kNcclFloat32, kNcclSum, (ncclComm_t)h->comm, s);
  if (r != 0) return nccl_fail("ncclAllReduce", r);
  return 0;
}
#endif

}  // namespace tcr
