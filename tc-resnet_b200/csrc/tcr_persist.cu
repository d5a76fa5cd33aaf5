// tcr_persist.cu — the whole training step (all layers forward, head, backward-data chain, weight gradients,
// gradient finalisation) as ONE persistent cooperative kernel.
//
// OPT-IN (TCR_PERSISTENT=1) and currently slower than the default multi-kernel path: kept as the
// starting point for a cooperative + cluster launch.  The kernel bodies of the multi-kernel path run as PHASES over virtual
// CTAs; a grid barrier replaces the kernel boundary and the BatchNorm statistics / BN-backward sums are finalised by a
// distributed phase (one warp per channel, lanes stride the per-CTA records, plain sums in a fixed order).
// The host records the phases with the same code that would launch the kernels one by one (h->rec != nullptr).
// Launched cooperatively (all CTAs co-resident: 2 per SM) so the spin barrier cannot deadlock.
#include <stdlib.h>

#include "tcr_bn.cuh"
#include "tcr_net.h"

namespace tcr {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Distributed finalize: one warp per channel; lanes stride the per-CTA partials in batches of kFinB independent loads
// (the unbatched version was a 7 us latency chain per phase), fixed order -> deterministic.
constexpr int kFinB = 8;
__device__ __forceinline__ void fin_fwd_phase(const FinFwd& F, int b, int nb) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int total = F.f[0].c + (F.nf > 1 ? F.f[1].c : 0);
  for (int item = b * wpb + warp; item < total; item += nb * wpb) {
    const int fi = item >= F.f[0].c ? 1 : 0;
    const int c = item - (fi ? F.f[0].c : 0);
    const BnFinalize& f = F.f[fi];
    float a1 = 0.f, a2 = 0.f;                              // per-lane fp32 partial sums (<= G/32 terms), fp64 across lanes
    for (int g0 = lane; g0 < F.G; g0 += 32 * kFinB) {
      float v1[kFinB], v2[kFinB];
#pragma unroll
      for (int j = 0; j < kFinB; ++j) {
        const int g = g0 + 32 * j;
        v1[j] = g < F.G ? __ldcg(f.fpart + ((size_t)g * f.c + c) * 2) : 0.f;
        v2[j] = g < F.G ? __ldcg(f.fpart + ((size_t)g * f.c + c) * 2 + 1) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < kFinB; ++j) { a1 += v1[j]; a2 += v2[j]; }
    }
    const double s1 = warp_sum_d((double)a1), s2 = warp_sum_d((double)a2);
    if (lane == 0) bn_table_write(f, c, s1, s2, (double)F.n * F.t_out, F.eps);
  }
}

__device__ __forceinline__ void fin_bwd_phase(const FinBwd& F, int b, int nb) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int n0 = 2 * F.f[0].c, n1 = F.nf > 1 ? 2 * F.f[1].c : 0;
  const int total = n0 + n1 + (F.loss_part ? 1 : 0);
  for (int item = b * wpb + warp; item < total; item += nb * wpb) {
    const float* src;
    size_t stride;
    float* dst;
    if (item < n0 + n1) {
      const int fi = item >= n0 ? 1 : 0;
      const int i = item - (fi ? n0 : 0);
      const BwdSumFinalize& f = F.f[fi];
      src = f.bpart + i;
      stride = (size_t)f.c * 2;
      dst = f.bsum + (i & 1) * f.c + (i >> 1);
    } else {
      src = F.loss_part;
      stride = 1;
      dst = F.loss_out;
    }
    double s = 0.0;
    for (int g0 = lane; g0 < F.G; g0 += 32 * kFinB) {
      float v[kFinB];
#pragma unroll
      for (int j = 0; j < kFinB; ++j) {
        const int g = g0 + 32 * j;
        v[j] = g < F.G ? __ldcg(src + (size_t)g * stride) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < kFinB; ++j) s += (double)v[j];
    }
    s = warp_sum_d(s);
    if (lane == 0) *dst = (float)s;
  }
}

__global__ void __launch_bounds__(kThreads, 2) step_kernel(const __grid_constant__ StepProgram P) {
  TCR_DYNAMIC_SMEM(smem_raw);
  MbarCtx mb{reinterpret_cast<uint64_t*>(smem_raw), 0u, false};
  if (threadIdx.x == 0) mbar_init(mb.bar, 1);
  __syncthreads();
  mb.ready = true;
  unsigned epoch = 0;
  const int b = (int)blockIdx.x, nb = (int)gridDim.x;
  float* scratch = reinterpret_cast<float*>(smem_raw) + 8;
  for (int ph = 0; ph < P.nphases; ++ph) {
    const Phase p = P.phase[ph];
    if (P.tl && threadIdx.x == 0 && ph < 64) {          // debug timeline: [phase][cta][start, work done]
      unsigned long long t;
#ifndef TCR_EMU
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
#else
      t = 0;
#endif
      P.tl[((size_t)ph * 512 + b) * 2] = (long long)t;
    }
    switch (p.kind) {
      case PH_TRANSPOSE:
        for (int64_t i = (int64_t)b * kThreads + threadIdx.x; i < P.wt.total; i += (int64_t)nb * kThreads) weight_transpose_body(P.wt, i);
        break;
      case PH_FWD:
        for (int vb = b; vb < p.nvb; vb += nb) {
          const FwdArgs& a = P.fwd[p.idx];
          if (p.k == 9) {
            if (p.wsmem) conv_fwd_body<9, true>(a, vb, p.nvb, smem_raw, mb);
            else conv_fwd_body<9, false>(a, vb, p.nvb, smem_raw, mb);
          } else {
            if (p.wsmem) conv_fwd_body<3, true>(a, vb, p.nvb, smem_raw, mb);
            else conv_fwd_body<3, false>(a, vb, p.nvb, smem_raw, mb);
          }
          __syncthreads();
        }
        break;
      case PH_FIN_FWD:
        fin_fwd_phase(P.finf[p.idx], b, nb);
        break;
      case PH_HEAD:
        for (int vb = b; vb < p.nvb; vb += nb) {
          head_body(P.head, vb, p.nvb, smem_raw);
          __syncthreads();
        }
        break;
      case PH_BWD:
        for (int vb = b; vb < p.nvb; vb += nb) {
          const BwdDataArgs& a = P.bwd[p.idx];
          if (p.wsmem) conv_bwd_data_body<9, true>(a, vb, p.nvb, smem_raw, mb);
          else conv_bwd_data_body<9, false>(a, vb, p.nvb, smem_raw, mb);
          __syncthreads();
        }
        break;
      case PH_FIN_BWD:
        fin_bwd_phase(P.finb[p.idx], b, nb);
        break;
      case PH_DW:
        for (int vb = b; vb < p.nvb; vb += nb) {
          dw_grouped_body(P.dw_layers, P.n_dw_layers, P.n, P.feat, nullptr, vb, smem_raw, BsumSrc{nullptr, 0, nullptr});
          __syncthreads();
        }
        break;
      case PH_GRAD:
        for (int vb = b; vb < p.nvb; vb += nb) {
          grad_finalize_body(P.grad, vb, scratch);
          __syncthreads();
        }
        break;
    }
    if (P.tl && ph < 64) {
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned long long t;
#ifndef TCR_EMU
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
#else
        t = 0;
#endif
        P.tl[((size_t)ph * 512 + b) * 2 + 1] = (long long)t;
      }
    }
    if (ph + 1 < P.nphases) grid_barrier(P.barrier, (unsigned)nb, epoch);
  }
}

// ------------------------------------------------------------------------------------------------
// host: recorder
// ------------------------------------------------------------------------------------------------
bool persist_enabled(tcr_handle* h) {
  if (h->persist < 0) {
    const char* e = getenv("TCR_PERSISTENT");
    h->persist = (e && e[0] == '1') ? 1 : 0;   // opt-in: the multi-kernel path is faster at present
#ifndef TCR_EMU
    int dev = 0, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) h->persist = 0;
#endif
  }
  return h->persist == 1;
}

void rec_begin(tcr_handle* h) {
  if (!h->rec) h->rec = new StepProgram();
  memset(h->rec, 0, sizeof(StepProgram));
  h->rec_smem = 1024;
}
void rec_abort(tcr_handle* h) {
  delete h->rec;
  h->rec = nullptr;
}
static void push_phase(tcr_handle* h, int kind, int idx, int nvb, int k, int wsm, size_t smem) {
  StepProgram& P = *h->rec;
  if (P.nphases >= kMaxPhases) return;
  P.phase[P.nphases++] = Phase{kind, idx, nvb, k, wsm};
  h->rec_smem = std::max(h->rec_smem, smem);
}
void rec_transpose(tcr_handle* h, const WtArgs& w) {
  h->rec->wt = w;
  push_phase(h, PH_TRANSPOSE, 0, 0, 0, 0, 0);
}
void rec_fwd(tcr_handle* h, const FwdArgs& a, int k, int wsm, int groups, size_t smem) {
  StepProgram& P = *h->rec;
  const int i = P.nfwd++;
  P.fwd[i] = a;
  push_phase(h, PH_FWD, i, groups, k, wsm, smem);
  FinFwd F;
  memset(&F, 0, sizeof(F));
  F.f[0] = a.fin;
  F.nf = 1;
  if (a.wd) { F.f[1] = a.find; F.nf = 2; }
  F.G = groups; F.U = a.U; F.n = a.n; F.t_out = a.t_out; F.eps = a.eps;
  P.finf[i] = F;
  push_phase(h, PH_FIN_FWD, i, 0, 0, 0, 0);
}
void rec_head(tcr_handle* h, const HeadArgs& a, int groups, size_t smem) {
  StepProgram& P = *h->rec;
  P.head = a;
  push_phase(h, PH_HEAD, 0, groups, 0, 0, smem);
  FinBwd F;
  memset(&F, 0, sizeof(F));
  F.f[0] = a.finb;
  F.nf = 1;
  if (a.ydn) { F.f[1] = a.find; F.nf = 2; }
  F.G = groups; F.loss_part = a.loss_part; F.loss_out = a.loss_out;
  const int i = P.nfinb++;
  P.finb[i] = F;
  push_phase(h, PH_FIN_BWD, i, 0, 0, 0, 0);
}
void rec_bwd(tcr_handle* h, const BwdDataArgs& a, int k, int wsm, int groups, size_t smem) {
  StepProgram& P = *h->rec;
  const int i = P.nbwd++;
  P.bwd[i] = a;
  push_phase(h, PH_BWD, i, groups, k, wsm, smem);
  FinBwd F;
  memset(&F, 0, sizeof(F));
  F.f[0] = a.finp;
  F.nf = 1;
  if (a.epi_kind == 2 && a.ypd) { F.f[1] = a.finpd; F.nf = 2; }
  F.G = groups;
  const int j = P.nfinb++;
  P.finb[j] = F;
  push_phase(h, PH_FIN_BWD, j, 0, 0, 0, 0);
}
void rec_dw(tcr_handle* h, int n, const float* feat) {
  StepProgram& P = *h->rec;
  P.dw_layers = h->d_dw_layers; P.n_dw_layers = h->n_dw_layers; P.n = n; P.feat = feat; P.tl = h->d_timeline;
  push_phase(h, PH_DW, 0, h->dw_ctas, 0, 0, h->dw_smem);
}
void rec_grad(tcr_handle* h, const GradArgs& g, int blocks) {
  h->rec->grad = g;
  push_phase(h, PH_GRAD, 0, blocks, 0, 0, 4096);
}

int rec_launch(tcr_handle* h, cudaStream_t s) {
  StepProgram& P = *h->rec;
  if (P.nfwd > kMaxFwdPh || P.nbwd > kMaxBwdPh || P.nphases >= kMaxPhases) {
    rec_abort(h);
    set_error("network too deep for the persistent step program");
    return TCR_ERR_UNSUPPORTED;
  }
  const size_t smem = h->rec_smem;
  auto kfn = step_kernel;
  int grid = 3;     // emulator: a few co-resident CTAs are enough to exercise every path
#ifndef TCR_EMU
  if (smem > h->persist_smem) {     // per handle (hence per device): the occupancy query below depends on it
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { rec_abort(h); return TCR_ERR_CUDA; }
    h->persist_smem = smem;
    h->persist_grid = 0;
  }
  if (!h->persist_grid) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kThreads, h->persist_smem) != cudaSuccess || per_sm < 1) {
      rec_abort(h);
      set_error("persistent step kernel does not fit on an SM");
      return TCR_ERR_CUDA;
    }
    h->persist_grid = sms * std::min(per_sm, 2);
  }
  grid = h->persist_grid;
#endif
  P.barrier = h->d_gridbar;
  P.tl = h->d_timeline;
  if (cudaMemsetAsync(h->d_gridbar, 0, sizeof(unsigned), s) != cudaSuccess) { rec_abort(h); return TCR_ERR_CUDA; }
  TCR_LAUNCH_COOP("step_persistent", kfn, dim3(grid), dim3(kThreads), smem, s, P);
  delete h->rec;
  h->rec = nullptr;
  return 0;
}

}  // namespace tcr
