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

void comm_destroy(tcr_handle* h) {
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)h->comm);
  h->comm = nullptr;
  h->world = 1;
  h->rank = 0;
}

int comm_allreduce_sum(tcr_handle* h, float* buf, int64_t count, cudaStream_t s) {
  ncclResult_t r = g_nccl.AllReduce(buf, buf, (size_t)count, kNcclFloat32, kNcclSum, (ncclComm_t)h->comm, s);
  if (r != 0) return nccl_fail("ncclAllReduce", r);
  return 0;
}
#endif

}  // namespace tcr
