// tcr_prof.cu — launch counter and optional per-kernel CUDA-event timing behind TCR_LAUNCH.
// bench.py uses it for `gpu_launches` and for the per-kernel durations of the roofline block; tests use the
// counter to prove that the CUDA path (not a fallback) ran.  Disabled profiling costs one branch per launch.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "tcr_net.h"

namespace tcr {

namespace {
struct Rec { std::string name; cudaEvent_t e0, e1; };
struct Profiler {
  bool enabled = false;
  unsigned long long launches = 0;
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  size_t pool_next = 0;
  std::map<std::string, std::pair<double, long>> acc;   // name -> (total ms, launches)
  std::vector<tcr_kernel_stat> out;
#ifndef TCR_EMU
  cudaEvent_t get() {
    if (pool_next == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[pool_next++];
  }
#endif
} g_prof;
}  // namespace

#ifndef TCR_EMU
// PDL only links the library's own kernels WITHIN one API call: the first launch of a call is an ordinary launch (fully
// ordered behind whatever the caller put on the stream), later ones carry the programmatic-serialization attribute.
static thread_local bool t_chain = false;
void pdl_chain_reset() { t_chain = false; }
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TCR_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  const bool use = v != 0 && t_chain;
  t_chain = true;
  return use;
}
void prof_begin(const char* name, cudaStream_t s) {
  ++g_prof.launches;
  if (!g_prof.enabled) return;
  Rec r{name, g_prof.get(), g_prof.get()};
  cudaEventRecord(r.e0, s);
  g_prof.recs.push_back(r);
}
void prof_end(cudaStream_t s) {
  if (!g_prof.enabled) return;
  cudaEventRecord(g_prof.recs.back().e1, s);
}
static void drain() {
  for (auto& r : g_prof.recs) {
    cudaEventSynchronize(r.e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    auto& a = g_prof.acc[r.name];
    a.first += ms;
    a.second += 1;
  }
  g_prof.recs.clear();
  g_prof.pool_next = 0;
}
#else
void pdl_chain_reset() {}
static void drain() {}
#endif

}  // namespace tcr

using namespace tcr;

extern "C" int tcr_profile_enable(int enable) {
  drain();
  g_prof.enabled = enable != 0;
  if (enable) g_prof.acc.clear();
  return TCR_OK;
}

extern "C" int tcr_profile_read(const tcr_kernel_stat** stats, int32_t* count) {
  if (!stats || !count) return TCR_ERR_INVALID;
  drain();
  g_prof.out.clear();
  for (auto& kv : g_prof.acc) {
    tcr_kernel_stat st;
    memset(&st, 0, sizeof(st));
    snprintf(st.name, sizeof(st.name), "%s", kv.first.c_str());
    st.total_ms = kv.second.first;
    st.launches = kv.second.second;
    g_prof.out.push_back(st);
  }
  *stats = g_prof.out.data();
  *count = (int32_t)g_prof.out.size();
  return TCR_OK;
}

extern "C" int tcr_launch_count(uint64_t* launches) {
  if (!launches) return TCR_ERR_INVALID;
  *launches = g_prof.launches;
  return TCR_OK;
}
