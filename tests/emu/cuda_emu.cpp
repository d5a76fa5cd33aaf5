// cuda_emu.cpp — fiber scheduler behind tests/emu/cuda_emu.h (TEST-ONLY, see header).
#include "cuda_emu.h"

namespace emu {

static State g_state;
State& st() { return g_state; }

namespace {
constexpr size_t kStackBytes = 256 * 1024;

struct Fiber {
  ucontext_t ctx;
  ThreadCtx tc;
  bool done = true;
  char* stack = nullptr;
};
struct WarpState {
  int count = 0;
  unsigned gen = 0;
  int alive = 0;
  uint64_t slots[32];
};

std::vector<Fiber> g_fibers;
std::vector<WarpState> g_warps;
ucontext_t g_sched;
const std::function<void()>* g_body = nullptr;
int g_nthreads = 0, g_alive = 0, g_cur = 0;
int g_bar_count = 0;
unsigned g_bar_gen = 0;
unsigned long long g_progress = 0;
std::vector<unsigned char> g_smem;

void yield() { swapcontext(&g_fibers[g_cur].ctx, &g_sched); }

void on_exit_thread() {
  Fiber& f = g_fibers[g_cur];
  f.done = true;
  ++g_progress;
  --g_alive;
  WarpState& w = g_warps[g_cur / 32];
  w.slots[g_cur % 32] = 0;
  --w.alive;
  if (w.alive > 0 && w.count == w.alive) { w.count = 0; ++w.gen; }
  if (g_alive > 0 && g_bar_count == g_alive) { g_bar_count = 0; ++g_bar_gen; }
}

void fiber_entry() {
  (*g_body)();
  on_exit_thread();
  swapcontext(&g_fibers[g_cur].ctx, &g_sched);
  abort();  // never resumed
}
}  // namespace

void syncthreads() {
  unsigned gen = g_bar_gen;
  ++g_progress;  // an arrival is progress (deadlock detector)
  if (++g_bar_count == g_alive) {
    g_bar_count = 0;
    ++g_bar_gen;
  } else {
    while (g_bar_gen == gen) yield();
  }
}

void syncwarp() {
  WarpState& w = g_warps[g_cur / 32];
  unsigned gen = w.gen;
  ++g_progress;
  if (++w.count == w.alive) {
    w.count = 0;
    ++w.gen;
  } else {
    while (w.gen == gen) yield();
  }
}

int lane_id() { return g_cur % 32; }

uint64_t warp_exchange(uint64_t v, int src_lane) {
  WarpState& w = g_warps[g_cur / 32];
  w.slots[g_cur % 32] = v;
  syncwarp();
  uint64_t r = w.slots[src_lane & 31];
  syncwarp();
  return r;
}

unsigned ballot(int pred) {
  WarpState& w = g_warps[g_cur / 32];
  w.slots[g_cur % 32] = pred ? 1 : 0;
  syncwarp();
  unsigned m = 0;
  for (int i = 0; i < 32; ++i)
    if (w.slots[i]) m |= 1u << i;
  syncwarp();
  return m;
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  const int n = (int)(block.x * block.y * block.z);
  if (n <= 0 || n > 1024) { fprintf(stderr, "emu: bad block size %d\n", n); abort(); }
  if ((int)g_fibers.size() < n) g_fibers.resize(n);
  for (int i = 0; i < n; ++i)
    if (!g_fibers[i].stack) g_fibers[i].stack = (char*)malloc(kStackBytes);
  g_warps.assign((n + 31) / 32, WarpState());
  g_smem.resize(smem_bytes + 2048);
  g_body = &body;
  g_state.bdim = block;
  g_state.gdim = grid;
  unsigned char* smem_base = (unsigned char*)(((uintptr_t)g_smem.data() + 1023) & ~(uintptr_t)1023);
  g_state.dyn_smem = smem_base;

  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_state.bid = uint3{bx, by, bz};
        memset(smem_base, 0xFF, smem_bytes);  // NaN poison
        g_nthreads = g_alive = n;
        g_bar_count = 0;
        for (auto& w : g_warps) { w.count = 0; w.alive = 0; memset(w.slots, 0, sizeof(w.slots)); }
        for (int i = 0; i < n; ++i) {
          Fiber& f = g_fibers[i];
          f.done = false;
          f.tc.tid = uint3{(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y), (unsigned)(i / (block.x * block.y))};
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack;
          f.ctx.uc_stack.ss_size = kStackBytes;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, fiber_entry, 0);
          g_warps[i / 32].alive++;
        }
        while (g_alive > 0) {
          unsigned long long before = g_progress;
          for (int i = 0; i < n; ++i) {
            if (g_fibers[i].done) continue;
            g_cur = i;
            g_state.cur = &g_fibers[i].tc;
            swapcontext(&g_sched, &g_fibers[i].ctx);
          }
          if (g_alive > 0 && g_progress == before) {
            fprintf(stderr, "emu: deadlock in block (%u,%u,%u): %d threads stuck at a barrier\n", bx, by, bz, g_alive);
            abort();
          }
        }
      }
  g_state.cur = nullptr;
}

}  // namespace emu
