// tcr_optim.cu — gradient finalisation, SGD-momentum update, BN moving averages, loss scalars.
//
// Replaces, per training step: the 12/18 L2Loss nodes + AddN (factory/audio_nets.py:175-182), the 32/50
// per-variable ApplyMomentum nodes (tf.train.MomentumOptimizer, helper/trainer.py:188-190: m <- mom*m + g;
// v <- v - lr*m) and the 20/32 AssignSub moving-average updates of slim.batch_norm (decay 0.997, UNBIASED
// batch variance, no zero-debias) with two flat multi-tensor kernels over the parameter buffer:
//   grad_finalize_kernel : g[p] = sum_r partial[r][p] (+ weight_decay * w[p]); per-block sum of w^2
//   update_kernel        : m <- mom*m + g/world ; w <- w - lr*m ; moving stats ; loss scalars
// The NCCL all-reduce of the flat gradient sits between the two when a communicator is attached.
#include "tcr_bn.cuh"
#include "tcr_net.h"

namespace tcr {

constexpr int kOptThreads = 256;

__device__ __forceinline__ int find_segment(const OptSegment* segs, int nsegs, int64_t p) {
  int lo = 0, hi = nsegs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (segs[mid].offset <= p) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ void grad_finalize_body(const GradArgs& a, const int vb, float* s_red) {
  const OptSegment* __restrict__ segs = a.segs;
  const int nsegs = a.nsegs, fc_seg = a.fc_seg, fc_R = a.fc_R;
  const int64_t total = a.total;
  const float* __restrict__ fc_part = a.fc_part;
  const float* __restrict__ params = a.params;
  const float weight_decay = a.weight_decay;
  float* __restrict__ grads = a.grads;
  float* __restrict__ l2part = a.l2part;
  const int64_t p = (int64_t)vb * kOptThreads + threadIdx.x;
  float w2 = 0.f;
  if (p < total) {
    const int si = find_segment(segs, nsegs, p);
    const OptSegment sg = segs[si];
    const int64_t i = p - sg.offset;
    const float* part = si == fc_seg ? fc_part : sg.part;
    const int R = si == fc_seg ? fc_R : sg.R;
    float g = 0.f;
    if (part) {                      // fixed summation order (deterministic); many loads in flight: the fc partials come one per
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // CTA of the resident forward kernel (148 records), and a
      int r = 0;                                                    // dependent round trip per 8 records made them the kernel's tail
      for (; r + 32 <= R; r += 32) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = part[(size_t)(r + j) * sg.numel + i];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j & 7] += v[j];
      }
      for (; r < R; r += 8) {        // last (partial) batches: predicated, still independent loads
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = r + j < R ? part[(size_t)(r + j) * sg.numel + i] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
      g = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    }
    if (sg.decay) {
      const float w = params[p];
      g = fmaf(weight_decay, w, g);
      w2 = w * w;
    } else {
      g *= a.bn_grad_scale;          // gamma / beta (the only segments without decay)
    }
    grads[p] = g;
  }
  w2 = warp_sum(w2);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = w2;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kOptThreads / 32; ++i) s += s_red[i];
    l2part[vb] = s;
  }
}
__global__ void __launch_bounds__(kOptThreads) grad_finalize_kernel(GradArgs a) {
  pdl_wait();
  __shared__ float s_red[kOptThreads / 32];
  grad_finalize_body(a, (int)blockIdx.x, s_red);
}

struct UpdateArgs {
  int64_t total;
  float* params; float* slots; const float* grads; float* moving;
  float lr, momentum, weight_decay, one_minus_decay, grad_scale;
  const MovingSegment* msegs; int nmsegs; int n;
  const float* l2part; int l2blocks;
  const float* ce_sum; int ce_count; float inv_n;      // cross-entropy records of the head launch
  float* losses;          // may be null
  float* grads_out;       // may be null: scaled gradient actually applied
  int apply;
  int param_blocks;
  // peer-memory exchange (world > 1 with tcr_comm_p2p_attach): every rank's gradient buffer and flag array as mapped here
  int p2p_world, p2p_rank;
  unsigned p2p_step;
  const float* peer_grads[8];
  unsigned* peer_flags[8];
};

// System-scope flag accesses for the cross-GPU arrival barrier.
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
#ifndef TCR_EMU
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
  *p = v;
#endif
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
#ifndef TCR_EMU
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#else
  return *p;
#endif
}

__global__ void __launch_bounds__(kOptThreads) update_kernel(UpdateArgs a) {
  pdl_wait();
  if (a.p2p_world > 1) {
    // Cross-GPU arrival barrier in peer memory.  The gradient of this rank is complete (previous kernel); block 0 announces it
    // in every rank's flag array, every CTA waits until all ranks have announced this step.  A rank cannot be more than one
    // step ahead of the slowest one, and the gradient buffers alternate with the step parity, so nobody overwrites what a
    // peer is still reading.
    if (threadIdx.x == 0) {
      if (blockIdx.x == 0) {
        __threadfence_system();
        for (int r = 0; r < a.p2p_world; ++r) st_release_sys(a.peer_flags[r] + a.p2p_rank, a.p2p_step);
      }
      const unsigned* mine = a.peer_flags[a.p2p_rank];
      for (int r = 0; r < a.p2p_world; ++r)
        while ((int)(ld_acquire_sys(mine + r) - a.p2p_step) < 0) __nanosleep(100);
    }
    __syncthreads();
  }
  const int64_t p = (int64_t)blockIdx.x * kOptThreads + threadIdx.x;
  if (p < a.total) {
    float gsum;
    if (a.p2p_world > 1) {
      gsum = 0.f;
      for (int r = 0; r < a.p2p_world; ++r) gsum += __ldcv(a.peer_grads[r] + p);     // fixed rank order: replicas stay identical
    } else {
      gsum = a.grads[p];
    }
    const float g = gsum * a.grad_scale;
    if (a.grads_out) a.grads_out[p] = g;
    if (a.apply) {
      const float m = fmaf(a.slots[p], a.momentum, g);     // accum = accum * momentum + grad
      a.slots[p] = m;
      a.params[p] = a.params[p] - a.lr * m;                // var -= lr * accum
    }
  }
  // extra CTAs past the parameter range: one per BN layer (moving averages) + one for the loss scalars, so no
  // serial tail hangs off block 0
  const int extra = (int)blockIdx.x - a.param_blocks;
  if (extra >= 0 && extra < a.nmsegs) {
    if (a.apply && a.moving) {
      const MovingSegment ms = a.msegs[extra];
      const float m_rows = (float)a.n * (float)ms.t_out;
      const float unb = m_rows > 1.f ? m_rows / (m_rows - 1.f) : 1.f;
      for (int c = threadIdx.x; c < ms.c; c += kOptThreads) {
        float mm = a.moving[ms.mm_off + c], mv = a.moving[ms.mv_off + c];
        mm -= (mm - ms.bnf[c]) * a.one_minus_decay;      // assign_moving_average, zero_debias=False
        mv -= (mv - ms.var[c] * unb) * a.one_minus_decay;
        a.moving[ms.mm_off + c] = mm;
        a.moving[ms.mv_off + c] = mv;
      }
    }
  } else if (extra == a.nmsegs && a.losses) {
    // every record is loaded by its own thread (one round trip), the two fixed-order sums run on two warps side by side
    __shared__ double s_l2[kOptThreads], s_ce[kOptThreads], s_tot[2];
    double l2 = 0.0, cs = 0.0;
    for (int i = threadIdx.x; i < a.l2blocks; i += kOptThreads) l2 += (double)a.l2part[i];
    for (int i = threadIdx.x; i < a.ce_count; i += kOptThreads) cs += (double)a.ce_sum[i];   // records of the head, <= 256: one per thread
    s_l2[threadIdx.x] = l2;
    s_ce[threadIdx.x] = cs;
    __syncthreads();
    if (threadIdx.x == 0 || threadIdx.x == 32) {
      const double* src = threadIdx.x == 0 ? s_l2 : s_ce;
      double tot = 0.0;
      for (int i = 0; i < kOptThreads; ++i) tot += src[i];
      s_tot[threadIdx.x >> 5] = tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const float model = (float)s_tot[1] * a.inv_n;
      a.losses[1] = model;
      a.losses[0] = model + a.weight_decay * (float)(0.5 * s_tot[0]);
    }
  }
}

// Loss scalars for a forward-only call: sum of w^2 over the decayed variables in one CTA.
__global__ void __launch_bounds__(1024) loss_only_kernel(const OptSegment* __restrict__ segs, int nsegs, const float* __restrict__ params,
                                                         float weight_decay, const float* ce_sum, int ce_count, float inv_n, float* losses) {
  pdl_wait();
  __shared__ float s_red[32];
  float w2 = 0.f;
  for (int si = 0; si < nsegs; ++si) {
    if (!segs[si].decay) continue;
    for (int64_t i = threadIdx.x; i < segs[si].numel; i += blockDim.x) {
      const float w = params[segs[si].offset + i];
      w2 = fmaf(w, w, w2);
    }
  }
  w2 = warp_sum(w2);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = w2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (unsigned i = 0; i < blockDim.x / 32; ++i) s += (double)s_red[i];
    double ce = 0.0;
    for (int i = 0; i < ce_count; ++i) ce += (double)ce_sum[i];
    const float model = (float)ce * inv_n;
    losses[1] = model;
    losses[0] = model + weight_decay * (float)(0.5 * s);
  }
}


int build_opt_segments(tcr_handle* h) {
  std::vector<OptSegment> segs;
  std::vector<MovingSegment> msegs;
  for (auto& cv : h->convs) {
    segs.push_back(OptSegment{cv.w_off, cv.wnumel(), cv.dwpart, cv.dw_R, 1});
    segs.push_back(OptSegment{cv.beta_off, cv.cout, cv.bsum, 1, 0});             // d beta  = sum dz
    segs.push_back(OptSegment{cv.gamma_off, cv.cout, cv.bsum + cv.cout, 1, 0});  // d gamma = sum dz * xhat
    msegs.push_back(MovingSegment{cv.mm_off, cv.mv_off, cv.cout, cv.bnf, cv.var, cv.t_out});
  }
  segs.push_back(OptSegment{h->fc_off, (int64_t)h->c_last * h->cfg.num_classes, h->d_dwfc_part, 1, 1});
  segs.push_back(OptSegment{h->fc2_off, (int64_t)h->c_last * 2, nullptr, 0, 1});   // dead "ranges" head: decay only
  h->n_segs = (int)segs.size();
  h->n_msegs = (int)msegs.size();
  void* p = nullptr;
  if (cudaMalloc(&p, segs.size() * sizeof(OptSegment)) != cudaSuccess) return TCR_ERR_CUDA;
  h->allocs.push_back(p);
  h->d_segs = (OptSegment*)p;
  if (cudaMemcpy(p, segs.data(), segs.size() * sizeof(OptSegment), cudaMemcpyHostToDevice) != cudaSuccess) return TCR_ERR_CUDA;
  if (cudaMalloc(&p, msegs.size() * sizeof(MovingSegment)) != cudaSuccess) return TCR_ERR_CUDA;
  h->allocs.push_back(p);
  h->d_msegs = (MovingSegment*)p;
  if (cudaMemcpy(p, msegs.data(), msegs.size() * sizeof(MovingSegment), cudaMemcpyHostToDevice) != cudaSuccess) return TCR_ERR_CUDA;
  return 0;
}

static int fc_segment(const tcr_handle* h) { return (int)h->convs.size() * 3; }

int net_update(tcr_handle* h, const float* feat, const tcr_step_args* a, cudaStream_t s) {
  const int blocks = (int)((h->n_train + kOptThreads - 1) / kOptThreads);
  if (blocks > 4096) { set_error("parameter count too large for the l2 partial buffer"); return TCR_ERR_UNSUPPORTED; }
  const bool p2p = h->p2p.attached && h->world > 1;
  if (p2p) ++h->p2p.step;
  float* grads = p2p ? h->p2p.grads + (size_t)(h->p2p.step & 1u) * h->n_train : h->d_grads;
  GradArgs ga{h->d_segs, h->n_segs, h->n_train, fc_segment(h), h->d_dwfc_part, h->fc_records, a->params, a->weight_decay,
              grads, h->d_l2part, sync_bn_on(h) ? 1.0f / (float)h->world : 1.0f};
  int l2_records = blocks;
  if (resident_mode(h) == 2) {      // backward chain + weight gradients + this reduction in one cooperative kernel (tcr_resident.cu)
    int rc = resident_backward(h, feat, a, grads, &l2_records, s);
    if (rc) return rc;
  } else {
    TCR_LAUNCH("grad_finalize", grad_finalize_kernel, dim3(blocks), dim3(kOptThreads), 0, s, ga);
  }
  if (!p2p && h->comm && h->world > 1) {
    int rc = comm_allreduce_sum(h, h->d_grads, h->n_train, s);
    if (rc) return rc;
  }
  UpdateArgs u;
  u.total = h->n_train;
  u.params = a->params; u.slots = a->slots; u.grads = grads; u.moving = a->moving;
  u.p2p_world = p2p ? h->world : 1; u.p2p_rank = h->rank; u.p2p_step = h->p2p.step;
  for (int r = 0; r < 8; ++r) {
    u.peer_grads[r] = p2p && r < h->world ? h->p2p.peer_grads[r] + (size_t)(h->p2p.step & 1u) * h->n_train : nullptr;
    u.peer_flags[r] = p2p && r < h->world ? h->p2p.peer_flags[r] : nullptr;
  }
  u.lr = a->learning_rate; u.momentum = a->momentum; u.weight_decay = a->weight_decay;
  u.one_minus_decay = (float)(1.0 - (double)h->cfg.bn_decay);
  u.grad_scale = 1.0f / (float)h->world;
  u.msegs = h->d_msegs; u.nmsegs = h->n_msegs; u.n = sync_bn_on(h) ? a->n * h->world : a->n;   // rows behind the batch variance
  u.l2part = h->d_l2part; u.l2blocks = l2_records;
  u.ce_sum = h->d_loss_part; u.ce_count = h->loss_gc; u.inv_n = 1.0f / (float)a->n;
  u.losses = a->losses; u.grads_out = a->grads; u.apply = a->apply_update ? 1 : 0;
  u.param_blocks = blocks;
  TCR_LAUNCH("update", update_kernel, dim3(blocks + h->n_msegs + 1), dim3(kOptThreads), 0, s, u);
  return 0;
}

int launch_loss_only(tcr_handle* h, const float* params, float weight_decay, int n, float* losses, cudaStream_t s) {
  TCR_LAUNCH("loss_only", loss_only_kernel, dim3(1), dim3(1024), 0, s, h->d_segs, h->n_segs, params, weight_decay, h->d_loss_part,
             h->loss_gc, 1.0f / (float)n, losses);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// fp32 FMA peak (compute-roofline denominator): 8 independent FMA chains per thread, all SMs busy.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float seed) {
  pdl_wait();
  float a0 = seed, a1 = seed + 1.f, a2 = seed + 2.f, a3 = seed + 3.f, a4 = seed + 4.f, a5 = seed + 5.f, a6 = seed + 6.f, a7 = seed + 7.f;
  const float b = 1.0000001f, c = 1e-7f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
      a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    }
  }
  const float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (r == 12345.678f) out[0] = r;   // keeps the chains alive without a store in practice
}

int measure_fp32_peak(tcr_handle* h, double* tflops, cudaStream_t s) {
#ifdef TCR_EMU
  (void)h; (void)s;
  *tflops = 0.0;
  return 0;
#else
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 8, iters = 4096;
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return TCR_ERR_CUDA;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, s);
    fma_peak_kernel<<<blocks, 256, 0, s>>>(h->d_l2part, iters, 1.0f);
    cudaEventRecord(e1, s);
    if (cudaEventSynchronize(e1) != cudaSuccess) return TCR_ERR_CUDA;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return 0;
#endif
}

}  // namespace tcr
