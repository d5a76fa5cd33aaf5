// tcr_dscnn.cu — DS-CNN (Hello-Edge) forward pass for sm_100a: the reference's 2-D-conv comparison model
// (audio_nets/ds_cnn.py:20-118, BASELINE.json config 5), inference mode.
//
//   dscnn_conv_kernel      first layer: kh x kw conv from the single-channel feature map (+bias, BN, ReLU)
//   dscnn_dsblock_kernel   depthwise 3x3 (+bias, BN, ReLU) -> pointwise 1x1 (+bias, BN, ReLU) fused per tile of output rows:
//                          the input rows (+halo) and both filter banks are staged in shared memory, the depthwise result
//                          never leaves the SM, the pointwise conv is a register-tiled 4 positions x 4 channels contraction
//   dscnn_head_kernel      global average pool + fully connected (+bias) + softmax
// BatchNorm here has no gamma (slim default scale=False) and uses the moving statistics; bias + BN fold into one
// per-channel (scale, shift) pair computed while staging.
#ifndef TCR_EMU
#include <cuda.h>   // CUtensorMap (types only: the encoder is looked up through the runtime)
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <initializer_list>

#include "tcr_bn.cuh"
#include "tcr_net.h"

namespace tcr {

struct DsLayerDev {
  int type;                 // 0 conv, 1 separable
  int cin, cout, kh, kw, sh, sw, hin, win, hout, wout, pt, pl;
  int64_t w, b, beta, mm, mv;                // conv / depthwise
  int64_t pw, pb, pbeta, pmm, pmv;           // pointwise (separable only)
};

constexpr int kDsMaxLayers = 8;
struct DsNet {
  int nlayers, n, classes;
  DsLayerDev layer[kDsMaxLayers];
  int64_t fcw, fcb;
};

__device__ __forceinline__ void fold_bn(const float* p, int64_t b, int64_t beta, int64_t mm, int64_t mv, int c, float eps,
                                        float* scale, float* shift) {
  const float rstd = 1.0f / sqrtf(p[mv + c] + eps);
  scale[c] = rstd;
  shift[c] = (p[b + c] - p[mm + c]) * rstd + p[beta + c];
}

// ---- first layer: cin == 1 ----
__global__ void __launch_bounds__(256) dscnn_conv_kernel(DsLayerDev L, const float* __restrict__ params, const float* __restrict__ in,
                                                         float* __restrict__ out, float eps) {
  pdl_wait();
  TCR_DYNAMIC_SMEM(smem_raw);
  float* smem = reinterpret_cast<float*>(smem_raw);
  const int hp = (L.hout - 1) * L.sh + L.kh, wp = (L.wout - 1) * L.sw + L.kw;      // padded input extent
  float* xs = smem;                                   // [hp][wp]
  float* ws = xs + ((hp * wp + 3) & ~3);              // [kh*kw][cout]
  float* sc = ws + L.kh * L.kw * L.cout;              // [cout]
  float* sf = sc + L.cout;
  const int n = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < hp * wp; i += blockDim.x) {
    const int h = i / wp - L.pt, w = i % wp - L.pl;
    xs[i] = (h >= 0 && h < L.hin && w >= 0 && w < L.win) ? in[((size_t)n * L.hin + h) * L.win + w] : 0.f;
  }
  for (int i = tid; i < L.kh * L.kw * L.cout; i += blockDim.x) ws[i] = params[L.w + i];
  for (int c = tid; c < L.cout; c += blockDim.x) fold_bn(params, L.b, L.beta, L.mm, L.mv, c, eps, sc, sf);
  __syncthreads();
  const int c4n = L.cout >> 2, npos = L.hout * L.wout;
  for (int task = tid; task < npos * c4n; task += blockDim.x) {
    const int c4 = task % c4n, pos = task / c4n;
    const int oh = pos / L.wout, ow = pos - oh * L.wout;
    const float* xb = xs + oh * L.sh * wp + ow * L.sw;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < L.kh; ++i)
      for (int j = 0; j < L.kw; ++j) {
        const float x = xb[i * wp + j];
        const float4 w = ld4(ws + (i * L.kw + j) * L.cout + 4 * c4);
        acc.x = fmaf(x, w.x, acc.x); acc.y = fmaf(x, w.y, acc.y); acc.z = fmaf(x, w.z, acc.z); acc.w = fmaf(x, w.w, acc.w);
      }
    const float4 s = ld4(sc + 4 * c4), t = ld4(sf + 4 * c4);
    const float4 r = make_float4(fmaxf(fmaf(acc.x, s.x, t.x), 0.f), fmaxf(fmaf(acc.y, s.y, t.y), 0.f),
                                 fmaxf(fmaf(acc.z, s.z, t.z), 0.f), fmaxf(fmaf(acc.w, s.w, t.w), 0.f));
    st4(out + ((size_t)n * npos + pos) * L.cout + 4 * c4, r);
  }
}

// ---- depthwise-separable block, one CTA per (utterance, chunk of RH output rows) ----
constexpr int kDsTM = 4;
__global__ void __launch_bounds__(256) dscnn_dsblock_kernel(DsLayerDev L, int RH, int COT, const float* __restrict__ params,
                                                            const float* __restrict__ in, float* __restrict__ out, float eps) {
  pdl_wait();
  TCR_DYNAMIC_SMEM(smem_raw);
  float* smem = reinterpret_cast<float*>(smem_raw);
  const int C = L.cin, CO = L.cout;
  const int n = blockIdx.y, h0 = blockIdx.x * RH, rh = imin(RH, L.hout - h0), tid = threadIdx.x;
  const int hin_t = (RH - 1) * L.sh + L.kh;                      // staged input rows (with halo)
  const int wp = (L.wout - 1) * L.sw + L.kw;                     // padded width
  float* xs = smem;                                              // [hin_t][wp][C]
  float* ds = xs + (size_t)hin_t * wp * C;                       // [RH*wout][C]  depthwise output
  float* pws = ds + (size_t)RH * L.wout * C;                     // [C][COT]      pointwise weights, a tile of COT output channels
  float* dws = pws + (size_t)C * COT;                            // [kh*kw][C]    depthwise weights   (COT = CO unless the bank exceeds shared memory: DS-CNN-L)
  float* sc1 = dws + L.kh * L.kw * C;                            // folded BN of the depthwise stage
  float* sf1 = sc1 + C;
  float* sc2 = sf1 + C;                                          // folded BN of the pointwise stage
  float* sf2 = sc2 + CO;
  const int c4n = C >> 2;
  {
    const RowWalk w = row_walk(tid, blockDim.x, c4n);
    const int rows = hin_t * wp;
    if (w.row < rows) {
      int r = w.row / wp, col = w.row - r * wp;
      for (int row = w.row; row < rows; row += w.rstep) {
        const int h = h0 * L.sh - L.pt + r, x = col - L.pl;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h >= 0 && h < L.hin && x >= 0 && x < L.win) v = ld4(in + (((size_t)n * L.hin + h) * L.win + x) * C + 4 * w.c4);
        st4(xs + (size_t)row * C + 4 * w.c4, v);
        col += w.rstep;
        while (col >= wp) { col -= wp; ++r; }
      }
    }
  }
  auto load_slice = [&](int co0, int cot) {                      // pw[ci][co0 .. co0 + cot) -> pws[ci][COT]
    const int q4 = cot >> 2;
    for (int i = tid; i < C * q4; i += blockDim.x) {
      const int ci = i / q4, q = i - ci * q4;
      st4(pws + (size_t)ci * COT + 4 * q, ldg4(params + L.pw + (size_t)ci * CO + co0 + 4 * q));
    }
  };
  load_slice(0, imin(COT, CO));
  for (int i = tid; i < L.kh * L.kw * C / 4; i += blockDim.x) st4(dws + 4 * i, ldg4(params + L.w + 4 * i));
  for (int c = tid; c < C; c += blockDim.x) fold_bn(params, L.b, L.beta, L.mm, L.mv, c, eps, sc1, sf1);
  for (int c = tid; c < CO; c += blockDim.x) fold_bn(params, L.pb, L.pbeta, L.pmm, L.pmv, c, eps, sc2, sf2);
  __syncthreads();
  // depthwise: thread owns 4 channels, walks output positions
  const int npos = rh * L.wout;
  {
    const RowWalk w = row_walk(tid, blockDim.x, c4n);
    if (w.row < npos) {
      const float4 s = ld4(sc1 + 4 * w.c4), t = ld4(sf1 + 4 * w.c4);
      int oh = w.row / L.wout, ow = w.row - oh * L.wout;
      for (int pos = w.row; pos < npos; pos += w.rstep) {
        const float* xb = xs + ((size_t)(oh * L.sh) * wp + ow * L.sw) * C + 4 * w.c4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = 0; i < L.kh; ++i)
          for (int j = 0; j < L.kw; ++j) {
            const float4 x = ld4(xb + ((size_t)i * wp + j) * C);
            const float4 k = ld4(dws + (i * L.kw + j) * C + 4 * w.c4);
            acc.x = fmaf(x.x, k.x, acc.x); acc.y = fmaf(x.y, k.y, acc.y); acc.z = fmaf(x.z, k.z, acc.z); acc.w = fmaf(x.w, k.w, acc.w);
          }
        st4(ds + (size_t)pos * C + 4 * w.c4,
            make_float4(fmaxf(fmaf(acc.x, s.x, t.x), 0.f), fmaxf(fmaf(acc.y, s.y, t.y), 0.f),
                        fmaxf(fmaf(acc.z, s.z, t.z), 0.f), fmaxf(fmaf(acc.w, s.w, t.w), 0.f)));
        ow += w.rstep;
        while (ow >= L.wout) { ow -= L.wout; ++oh; }
      }
    }
  }
  __syncthreads();
  // pointwise: tasks of kDsTM positions x 4 output channels, positions of a task are NRT apart; one pass per tile of COT channels
  const int NRT = (npos + kDsTM - 1) / kDsTM;
  const size_t gbase = ((size_t)n * L.hout + h0) * L.wout;
  for (int co0 = 0; co0 < CO; co0 += COT) {
    const int cot = imin(COT, CO - co0), ncg = cot >> 2;
    if (co0 > 0) {
      __syncthreads();                                           // everybody is done with the previous slice
      load_slice(co0, cot);
      __syncthreads();
    }
    for (int task = tid; task < NRT * ncg; task += blockDim.x) {
      const int cg = task % ncg, rt = task / ncg;
      const float* xr[kDsTM];
#pragma unroll
      for (int i = 0; i < kDsTM; ++i) xr[i] = ds + (size_t)imin(rt + i * NRT, npos - 1) * C;
      float4 acc[kDsTM];
#pragma unroll
      for (int i = 0; i < kDsTM; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* wk = pws + 4 * cg;
#pragma unroll 2
      for (int ci = 0; ci < C; ci += 4) {
        const float4 w0 = ld4(wk + (ci + 0) * COT), w1 = ld4(wk + (ci + 1) * COT), w2 = ld4(wk + (ci + 2) * COT), w3 = ld4(wk + (ci + 3) * COT);
#pragma unroll
        for (int i = 0; i < kDsTM; ++i) {
          const float4 x = ld4(xr[i] + ci);
          acc[i].x = fmaf(x.x, w0.x, acc[i].x); acc[i].y = fmaf(x.x, w0.y, acc[i].y); acc[i].z = fmaf(x.x, w0.z, acc[i].z); acc[i].w = fmaf(x.x, w0.w, acc[i].w);
          acc[i].x = fmaf(x.y, w1.x, acc[i].x); acc[i].y = fmaf(x.y, w1.y, acc[i].y); acc[i].z = fmaf(x.y, w1.z, acc[i].z); acc[i].w = fmaf(x.y, w1.w, acc[i].w);
          acc[i].x = fmaf(x.z, w2.x, acc[i].x); acc[i].y = fmaf(x.z, w2.y, acc[i].y); acc[i].z = fmaf(x.z, w2.z, acc[i].z); acc[i].w = fmaf(x.z, w2.w, acc[i].w);
          acc[i].x = fmaf(x.w, w3.x, acc[i].x); acc[i].y = fmaf(x.w, w3.y, acc[i].y); acc[i].z = fmaf(x.w, w3.z, acc[i].z); acc[i].w = fmaf(x.w, w3.w, acc[i].w);
        }
      }
      const float4 s = ld4(sc2 + co0 + 4 * cg), t = ld4(sf2 + co0 + 4 * cg);
#pragma unroll
      for (int i = 0; i < kDsTM; ++i) {
        const int pos = rt + i * NRT;
        if (pos < npos)
          st4(out + (gbase + pos) * CO + co0 + 4 * cg,
              make_float4(fmaxf(fmaf(acc[i].x, s.x, t.x), 0.f), fmaxf(fmaf(acc[i].y, s.y, t.y), 0.f),
                          fmaxf(fmaf(acc[i].z, s.z, t.z), 0.f), fmaxf(fmaf(acc[i].w, s.w, t.w), 0.f)));
      }
    }
  }
}

// ---- depthwise-separable block with the pointwise 1x1 conv on the 5th-generation tensor cores (tcgen05 / UMMA) ----
// The pointwise conv IS a dense [positions, C] x [C, CO] contraction (audio_nets/ds_cnn.py:56-59), 77 % of DS-CNN-S's FLOPs.
// Per CTA: 128 accumulator rows (RH output rows x wout positions, RH * wout <= 128) x CO columns in TMEM.
//   * single TF32 products miss the 1e-4 logit bound (6.7e-4 relative on a K = 64 dot product, tools/ubench/umma_tf32.cu), so every
//     operand is split x = hi + lo with hi = x truncated to TF32 (low 13 mantissa bits cleared, lo = x - hi exact) and three MMAs
//     accumulate lo*hi + hi*lo + hi*hi in fp32: 4e-7 relative, the same as the FMA kernel;
//   * operands sit in shared memory in UMMA's canonical K-major no-swizzle layout, element (row, k) at float offset
//     (k / 4) * LBO + row * 4 + k % 4: an 8-row core matrix is 128 contiguous bytes, 8-row groups are 128 B apart (SBO) and the
//     K-chunk stride LBO is padded by 16 B so that neither the depthwise stage's float4 stores nor the filter staging conflict;
//   * one thread issues the 3 * C/8 tcgen05.mma (M = 128, N = CO, K = 8) and commits them to an mbarrier; all 8 warps then read
//     their TMEM lane quarter (warp % 4) and column half (warp / 4) with tcgen05.ld, apply the folded BatchNorm + ReLU, store.
// The input rows are staged channel-chunk-major ([C/4][rows*wp][4]) so that lanes walking positions read conflict-free float4s.
#ifndef TCR_EMU
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);                  // start address >> 4, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;         // leading (K-chunk) byte offset >> 4, bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;         // stride (8-row group) byte offset >> 4, bits [32,46)
  d |= (uint64_t)1 << 46;                                   // descriptor version 1 (sm_100); layout type 0 = no swizzle
  return d;
}
__device__ __forceinline__ void tf32_split(float4 v, float4& hi, float4& lo) {
  hi = make_float4(__uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u), __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u),
                   __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u), __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
  lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
}

__device__ __forceinline__ void cp_async16_zfill(float* dst_smem, const float* src, bool valid) {
  const uint32_t bytes = valid ? 16u : 0u;                        // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes) : "memory");
}

// Persistent: a CTA keeps the split filter bank, the TMEM allocation and the per-channel tables for all its tiles
// (tile = RH output rows of one utterance), and the next tile's input rows travel (cp.async) while the current tile is computed.
// Specialised for the 3x3 / stride-1 depthwise stage (every separable block of DS-CNN-S, audio_nets/ds_cnn.py:20-26):
//   * depthwise: a thread owns (4 channels, one output column) and slides down the rows, so every input float4 is read once
//     per thread (3 loads per output instead of 9) and the 9 taps stay in registers;
//   * two TMEM accumulators: the MMAs of tile i run while the epilogue of tile i-1 drains the other one;
//   * the UMMA descriptors are formed once; a k-step only adds its byte offset to the start-address field.
constexpr int kTcThreads = 512;
template <int C, int CO>
__global__ void __launch_bounds__(kTcThreads, 1) dscnn_dsblock_tc_kernel(DsLayerDev L, int RH, int n_utt, const float* __restrict__ params,
                                                                  const float* __restrict__ in, float* __restrict__ out, float eps) {
  TCR_DYNAMIC_SMEM(smem_raw);
  __shared__ uint64_t mma_bar[2];
  __shared__ uint32_t tmem_base_s;
  float* smem = reinterpret_cast<float*>(smem_raw);
  constexpr int C4 = C / 4, KS = C / 8;
  constexpr int TMEM_COLS = 2 * CO <= 32 ? 32 : (2 * CO <= 64 ? 64 : (2 * CO <= 128 ? 128 : 256));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hin_t = RH + 2;                                       // 3x3, stride 1
  const int wp = L.wout + 2;
  const int XP = hin_t * wp * 4 + 4;                              // floats per channel chunk of the input tile (+16 B against conflicts)
  constexpr int LBA = 128 * 4 + 4, LBB = CO * 4 + 4;              // K-chunk strides of the A / B operand tiles (floats)
  float* xs0 = smem;                                              // [2][C4][XP]
  float* a_hi = xs0 + (size_t)2 * C4 * XP;                        // [C4][LBA]
  float* a_lo = a_hi + (size_t)C4 * LBA;
  float* b_hi = a_lo + (size_t)C4 * LBA;                          // [C4][LBB]
  float* b_lo = b_hi + (size_t)C4 * LBB;
  float* dws = b_lo + (size_t)C4 * LBB;                           // [9][C]
  float* sc1 = dws + 9 * C;
  float* sf1 = sc1 + C;
  float* sc2 = sf1 + C;
  float* sf2 = sc2 + CO;
  const int tpu = (L.hout + RH - 1) / RH;                         // tiles per utterance
  const int ntiles = tpu * n_utt;
  if (tid == 0) { mbar_init(&mma_bar[0], 1); mbar_init(&mma_bar[1], 1); }
  if (warp == 0) {                                               // TMEM: two accumulators of 128 lanes x CO fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // pointwise filters pw[ci][co] -> B operand [n = co][k = ci], hi / lo parts (params are caller-owned: independent of the producer)
  for (int i0 = tid; i0 < C * CO; i0 += kTcThreads * 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = i0 + kTcThreads * j < C * CO ? __ldg(params + L.pw + i0 + kTcThreads * j) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + kTcThreads * j;
      if (i < C * CO) {
        const int ci = i / CO, co = i - ci * CO;
        const float hi = __uint_as_float(__float_as_uint(v[j]) & 0xFFFFE000u);
        const int o = (ci >> 2) * LBB + co * 4 + (ci & 3);
        b_hi[o] = hi;
        b_lo[o] = v[j] - hi;
      }
    }
  }
  for (int i = tid; i < 9 * C / 4; i += kTcThreads) st4(dws + 4 * i, ldg4(params + L.w + 4 * i));
  for (int c = tid; c < C; c += kTcThreads) fold_bn(params, L.b, L.beta, L.mm, L.mv, c, eps, sc1, sf1);
  for (int c = tid; c < CO; c += kTcThreads) fold_bn(params, L.pb, L.pbeta, L.pmm, L.pmv, c, eps, sc2, sf2);
  pdl_wait();
  // input rows (+halo) of a tile, channel-chunk-major, asynchronously.  A thread copies the same (row, column) cells of every
  // tile (cell = tid / C4 + j * (threads / C4)), so the walk is worked out once: cell -> (tile row, column validity, element offset).
  constexpr int kCells = 8, kRowStep = kTcThreads / C4;
  const int f_c4 = tid % C4, f_row0 = tid / C4, f_rows = hin_t * wp;
  int f_pk[kCells];
#pragma unroll
  for (int j = 0; j < kCells; ++j) {
    const int row = f_row0 + j * kRowStep;
    const int r = row / wp, col = row - r * wp, x = col - 1;
    const int okx = (row < f_rows && x >= 0 && x < L.win) ? 1 : 0;
    f_pk[j] = ((r * L.win + x + 1) << 9) | (r << 1) | okx;         // position offset + 1 (>= 0), tile row, column-valid bit
  }
  auto fetch_tile = [&](int tile, float* xs) {
    const int n = tile / tpu, h0 = (tile - n * tpu) * RH;
    const float* base = in + (((int64_t)n * L.hin + (h0 - 1)) * L.win - 1) * C + 4 * f_c4;      // never dereferenced out of range
    float* dst = xs + (size_t)f_c4 * XP + 4 * f_row0;
#pragma unroll
    for (int j = 0; j < kCells; ++j) {
      if (f_row0 + j * kRowStep < f_rows) {
        const int pk = f_pk[j], h = h0 - 1 + ((pk >> 1) & 0xFF);
        const bool ok = (pk & 1) && h >= 0 && h < L.hin;
        cp_async16_zfill(dst + 4 * j * kRowStep, ok ? base + (int64_t)(pk >> 9) * C : in, ok);
      }
    }
  };
  // epilogue of one tile: accumulator row m = 32 * (warp % 4) + lane, columns [(warp / 4) * CO / NG, + CO / NG), batches of 16
  auto epilogue = [&](uint32_t tacc, int tile_e) {
    const int n = tile_e / tpu, h0 = (tile_e - n * tpu) * RH, npos = imin(RH, L.hout - h0) * L.wout;
    constexpr int NG = kTcThreads / 128;                           // column groups: every warp reads its TMEM lane quarter
    const int m = 32 * (warp & 3) + lane, cbase = (warp >> 2) * (CO / NG);
    const size_t gbase = ((size_t)n * L.hout + h0) * L.wout;
#pragma unroll
    for (int c0 = 0; c0 < CO / NG; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tacc + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(cbase + c0);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (m < npos) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const int co = cbase + c0 + j;
          const float4 s = ld4(sc2 + co), t = ld4(sf2 + co);
          st4(out + (gbase + m) * CO + co,
              make_float4(fmaxf(fmaf(__uint_as_float(r[j]), s.x, t.x), 0.f), fmaxf(fmaf(__uint_as_float(r[j + 1]), s.y, t.y), 0.f),
                          fmaxf(fmaf(__uint_as_float(r[j + 2]), s.z, t.z), 0.f), fmaxf(fmaf(__uint_as_float(r[j + 3]), s.w, t.w), 0.f)));
        }
      }
    }
  };
  int tile = blockIdx.x, cur = 0, it = 0, prev_tile = -1;
  uint32_t parity[2] = {0u, 0u};
  if (tile < ntiles) fetch_tile(tile, xs0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // instruction descriptor: D fp32 (bits 4-5 = 1), A and B TF32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CO >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t da_hi = umma_desc(smem_u32(a_hi), LBA * 4u, 128u), da_lo = umma_desc(smem_u32(a_lo), LBA * 4u, 128u);
  const uint64_t db_hi = umma_desc(smem_u32(b_hi), LBB * 4u, 128u), db_lo = umma_desc(smem_u32(b_lo), LBB * 4u, 128u);
  for (; tile < ntiles; tile += gridDim.x, cur ^= 1, ++it) {
    const int n = tile / tpu, h0 = (tile - n * tpu) * RH, rh = imin(RH, L.hout - h0);
    (void)n;
    float* xs = xs0 + (size_t)cur * C4 * XP;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                                             // this tile's rows have landed
    if (tile + (int)gridDim.x < ntiles) fetch_tile(tile + gridDim.x, xs0 + (size_t)(cur ^ 1) * C4 * XP);
    // the previous tile's MMAs read the A tiles: they must be done before the depthwise stage overwrites them
    if (prev_tile >= 0) {
      mbar_wait(&mma_bar[(it - 1) & 1], parity[(it - 1) & 1]);
      parity[(it - 1) & 1] ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // depthwise 3x3 + folded BN + ReLU -> A operand (hi / lo): thread = (channel chunk, output column), sliding down the rows
    // row segments keep every thread busy: (segment, channel chunk, column) tasks, a segment re-reads its two halo rows
    const int ntask0 = C4 * L.wout;
    const int nseg = imax(1, imin(rh, kTcThreads / ntask0)), srows = (rh + nseg - 1) / nseg;
    for (int task = tid; task < ntask0 * nseg; task += kTcThreads) {
      const int seg = task / ntask0, t0 = task - seg * ntask0;
      const int c4 = t0 / L.wout, ow = t0 - c4 * L.wout;
      const int rb = seg * srows, re = imin(rh, rb + srows);       // output rows [rb, re) of the tile
      if (rb >= re) continue;
      float4 k[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) k[j] = ld4(dws + j * C + 4 * c4);
      const float4 s = ld4(sc1 + 4 * c4), t = ld4(sf1 + 4 * c4);
      const float* xb = xs + (size_t)c4 * XP + 4 * ow;
      float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0, acc2 = acc0;      // outputs r-2, r-1, r while input row r arrives
      for (int r = rb; r < re + 2; ++r) {
        const float4 x0 = ld4(xb + 4 * (r * wp)), x1 = ld4(xb + 4 * (r * wp + 1)), x2 = ld4(xb + 4 * (r * wp + 2));
#define TCR_DW3(acc, kr)                                                                                                                   \
  acc.x = fmaf(x0.x, k[3 * (kr)].x, fmaf(x1.x, k[3 * (kr) + 1].x, fmaf(x2.x, k[3 * (kr) + 2].x, acc.x)));                                  \
  acc.y = fmaf(x0.y, k[3 * (kr)].y, fmaf(x1.y, k[3 * (kr) + 1].y, fmaf(x2.y, k[3 * (kr) + 2].y, acc.y)));                                  \
  acc.z = fmaf(x0.z, k[3 * (kr)].z, fmaf(x1.z, k[3 * (kr) + 1].z, fmaf(x2.z, k[3 * (kr) + 2].z, acc.z)));                                  \
  acc.w = fmaf(x0.w, k[3 * (kr)].w, fmaf(x1.w, k[3 * (kr) + 1].w, fmaf(x2.w, k[3 * (kr) + 2].w, acc.w)));
        TCR_DW3(acc0, 2)              // input row r is the bottom tap row of output r-2, the middle of r-1, the top of r
        TCR_DW3(acc1, 1)
        TCR_DW3(acc2, 0)
#undef TCR_DW3
        if (r >= rb + 2) {
          float4 hi, lo;
          tf32_split(make_float4(fmaxf(fmaf(acc0.x, s.x, t.x), 0.f), fmaxf(fmaf(acc0.y, s.y, t.y), 0.f), fmaxf(fmaf(acc0.z, s.z, t.z), 0.f),
                                 fmaxf(fmaf(acc0.w, s.w, t.w), 0.f)), hi, lo);
          const int m = (r - 2) * L.wout + ow;
          st4(a_hi + (size_t)c4 * LBA + 4 * m, hi);
          st4(a_lo + (size_t)c4 * LBA + 4 * m, lo);
        }
        acc0 = acc1; acc1 = acc2; acc2 = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // accumulator rows past rh * wout are never stored, whatever the A rows there hold (each output row depends on its own A row only)
    fence_proxy_async();                                         // generic-proxy stores of the operands -> visible to the async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      const uint32_t tacc = tmem + (uint32_t)((it & 1) * CO);
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {                     // lo*hi, hi*lo, hi*hi: the small terms first
        const uint64_t da0 = pass == 0 ? da_lo : da_hi, db0 = pass == 1 ? db_lo : db_hi;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {                        // one MMA covers K = 8 = two 4-wide chunks
          const uint64_t da = da0 + (uint64_t)((2 * ks * LBA * 4) >> 4), db = db0 + (uint64_t)((2 * ks * LBB * 4) >> 4);
          const uint32_t accumulate = (pass | ks) ? 1u : 0u;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tacc), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_bar[it & 1])) : "memory");
    }
    // while these MMAs run: drain the previous tile's accumulator (its MMAs were waited for above)
    if (prev_tile >= 0) epilogue(tmem + (uint32_t)(((it - 1) & 1) * CO), prev_tile);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    prev_tile = tile;
  }
  if (prev_tile >= 0) {
    mbar_wait(&mma_bar[(it - 1) & 1], parity[(it - 1) & 1]);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    epilogue(tmem + (uint32_t)(((it - 1) & 1) * CO), prev_tile);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
}

// Warp-specialised version of the block above (the default): the stages of a tile run on different warps and meet on mbarriers,
// so the input fetch, the depthwise stage, the MMAs and the epilogue of neighbouring tiles overlap instead of taking turns.
//   warps 0-9   producers: TMA fetch of the tile two ahead (cp.async.bulk.tensor into the input buffer they have just left) and
//               the depthwise stage, input tile -> A operand (hi / lo); both double-buffered
//   warp  10    one elected thread issues the 3 * C/8 tcgen05.mma of a tile and commits them to `mma_done`
//   warps 12-15 epilogue (TMEM lane quarter = warp % 4, all CO columns)
// Barriers (index = tile parity): in_full (TMA complete_tx), a_full (producers), mma_done (tcgen05.commit),
// tmem_empty (epilogue warps).  A rows are packed to the tile's own row count (rounded to 8): the M = 128 MMA then reads a few
// rows of the next K chunk as rows >= npos, whose accumulator rows are never stored.
constexpr int kWsProducerWarps = 10, kWsProducers = 32 * kWsProducerWarps, kWsFetchers = 128;
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {       // arrives when this thread's earlier cp.asyncs have landed
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// The MMA and epilogue roles of the warp-specialised kernels (shared by the separable block and the first conv).
// MMA role, one thread: per tile 3 * K4/2 tcgen05.mma (M = 128, N = CO, K = 8) on A buffer it % 2 -> TMEM accumulator it % 2.
template <int CO>
__device__ __forceinline__ void ws_mma_role(int ntiles, int K4, int LBA, const float* a0, const float* b_hi, const float* b_lo, uint32_t tmem,
                                            uint64_t* a_full, uint64_t* tmem_empty, uint64_t* mma_done) {
  constexpr int LBB = CO * 4 + 4;
  // instruction descriptor: D fp32 (bits 4-5 = 1), A and B TF32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CO >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t db_hi = umma_desc(smem_u32(b_hi), LBB * 4u, 128u), db_lo = umma_desc(smem_u32(b_lo), LBB * 4u, 128u);
  const int KS = K4 >> 1;
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int b = it & 1, k = it >> 1;
    const float* a_hi = a0 + (size_t)b * 2 * K4 * LBA;
    const uint64_t da_hi = umma_desc(smem_u32(a_hi), (uint32_t)LBA * 4u, 128u);
    const uint64_t da_lo = umma_desc(smem_u32(a_hi + (size_t)K4 * LBA), (uint32_t)LBA * 4u, 128u);
    mbar_wait(&a_full[b], (uint32_t)(k & 1));
    if (it >= 2) mbar_wait(&tmem_empty[b], (uint32_t)((k - 1) & 1));              // the epilogue of tile it-2 has drained this accumulator
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tacc = tmem + (uint32_t)(b * CO);
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {                       // lo*hi, hi*lo, hi*hi: the small terms first
      const uint64_t da0 = pass == 0 ? da_lo : da_hi, db0 = pass == 1 ? db_lo : db_hi;
#pragma unroll 4
      for (int ks = 0; ks < KS; ++ks) {                          // one MMA covers K = 8 = two 4-wide chunks
        const uint64_t da = da0 + (uint64_t)((2 * ks * LBA * 4) >> 4), db = db0 + (uint64_t)((2 * ks * LBB * 4) >> 4);
        const uint32_t accumulate = (pass | ks) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tacc), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_done[b])) : "memory");
  }
}
// Epilogue role, 4 warps: warp quarter q reads TMEM lanes [32 q, 32 q + 32) = accumulator rows, all CO columns in batches of 16,
// applies the folded BatchNorm + ReLU and stores the NHWC row; then hands the accumulator back (tmem_empty).
template <int CO>
__device__ __forceinline__ void ws_epilogue_role(const DsLayerDev& L, int RH, int tpu, int ntiles, uint32_t tmem, const float* sc, const float* sf,
                                                 float* __restrict__ out, uint64_t* mma_done, uint64_t* tmem_empty, int q, int lane) {
  const int m = 32 * q + lane;
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int b = it & 1, k = it >> 1;
    mbar_wait(&mma_done[b], (uint32_t)(k & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int n = tile / tpu, h0 = (tile - n * tpu) * RH, npos = imin(RH, L.hout - h0) * L.wout;
    const size_t gbase = ((size_t)n * L.hout + h0) * L.wout;
    const uint32_t tacc = tmem + (uint32_t)(b * CO) + ((uint32_t)(32 * q) << 16);
#pragma unroll
    for (int c0 = 0; c0 < CO; c0 += 16) {
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(tacc + (uint32_t)c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (m < npos) {                                             // a lane owns a whole NHWC row: 32-byte stores (one full sector each)
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          const int co = c0 + j;
          const float4 s0 = ld4(sc + co), t0 = ld4(sf + co), s1 = ld4(sc + co + 4), t1 = ld4(sf + co + 4);
          asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out + (gbase + m) * CO + co),
                       "f"(fmaxf(fmaf(__uint_as_float(r[j]), s0.x, t0.x), 0.f)), "f"(fmaxf(fmaf(__uint_as_float(r[j + 1]), s0.y, t0.y), 0.f)),
                       "f"(fmaxf(fmaf(__uint_as_float(r[j + 2]), s0.z, t0.z), 0.f)), "f"(fmaxf(fmaf(__uint_as_float(r[j + 3]), s0.w, t0.w), 0.f)),
                       "f"(fmaxf(fmaf(__uint_as_float(r[j + 4]), s1.x, t1.x), 0.f)), "f"(fmaxf(fmaf(__uint_as_float(r[j + 5]), s1.y, t1.y), 0.f)),
                       "f"(fmaxf(fmaf(__uint_as_float(r[j + 6]), s1.z, t1.z), 0.f)), "f"(fmaxf(fmaf(__uint_as_float(r[j + 7]), s1.w, t1.w), 0.f))
                       : "memory");
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive(&tmem_empty[b]);
  }
}

template <int C, int CO>
__global__ void __launch_bounds__(kTcThreads, 1) dscnn_dsblock_ws_kernel(const __grid_constant__ CUtensorMap in_map, DsLayerDev L, int RH, int n_utt,
                                                                         const float* __restrict__ params, float* __restrict__ out, float eps) {
  TCR_DYNAMIC_SMEM(smem_raw);
  __shared__ uint64_t in_full[2], a_full[2], mma_done[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_s;
  float* smem = reinterpret_cast<float*>(smem_raw);
  constexpr int C4 = C / 4, KS = C / 8;
  constexpr int TMEM_COLS = 2 * CO <= 32 ? 32 : (2 * CO <= 64 ? 64 : (2 * CO <= 128 ? 128 : 256));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hin_t = RH + 2, wp = L.wout + 2;
  constexpr int CH = C / 2, CH4 = CH / 4;                         // channels per TMA box: 32 fp32 = one 128-byte swizzle row
  static_assert(CH * 4 == 128, "the input halves are 128-byte rows (SWIZZLE_128B)");
  const int XH = ((hin_t * wp * CH * 4 + 1023) & ~1023) / 4;      // floats per half tile, a multiple of the 1024-byte swizzle atom
  const int LBA = ((RH * L.wout + 7) & ~7) * 4 + 4;               // K-chunk stride of the A tiles (floats): the tile's rows, +16 B
  constexpr int LBB = CO * 4 + 4;
  float* xs0 = smem;                                              // [2 buffers][2 channel halves][XH], TMA destination (1024-aligned)
  float* a0 = xs0 + (size_t)4 * XH;                               // [2][hi, lo][C4][LBA]
  float* b_hi = a0 + (size_t)4 * C4 * LBA;                        // [C4][LBB]  (also absorbs the last A chunk's over-read)
  float* b_lo = b_hi + (size_t)C4 * LBB;
  float* dws = b_lo + (size_t)C4 * LBB;                           // [9][C]
  float* sc1 = dws + 9 * C;
  float* sf1 = sc1 + C;
  float* sc2 = sf1 + C;
  float* sf2 = sc2 + CO;
  const int tpu = (L.hout + RH - 1) / RH;
  const int ntiles = tpu * n_utt;
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&in_full[b], 1);
      mbar_init(&a_full[b], kWsProducers);
      mbar_init(&mma_done[b], 1);
      mbar_init(&tmem_empty[b], kWsFetchers);
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i0 = tid; i0 < C * CO; i0 += kTcThreads * 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = i0 + kTcThreads * j < C * CO ? __ldg(params + L.pw + i0 + kTcThreads * j) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + kTcThreads * j;
      if (i < C * CO) {
        const int ci = i / CO, co = i - ci * CO;
        const float hi = __uint_as_float(__float_as_uint(v[j]) & 0xFFFFE000u);
        const int o = (ci >> 2) * LBB + co * 4 + (ci & 3);
        b_hi[o] = hi;
        b_lo[o] = v[j] - hi;
      }
    }
  }
  for (int i = tid; i < 9 * C / 4; i += kTcThreads) st4(dws + 4 * i, ldg4(params + L.w + 4 * i));
  for (int c = tid; c < C; c += kTcThreads) fold_bn(params, L.b, L.beta, L.mm, L.mv, c, eps, sc1, sf1);
  for (int c = tid; c < CO; c += kTcThreads) fold_bn(params, L.pb, L.pbeta, L.pmm, L.pmv, c, eps, sc2, sf2);
  pdl_wait();
  fence_proxy_async();                                           // the B operand was written through the generic proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (warp < kWsProducerWarps) {
    // ---- input fetch (TMA) + depthwise 3x3 + folded BN + ReLU -> A operand ----
    // One thread issues two 4-D tensor loads per tile (channel halves; box = 32 channels x (wout + 2) columns x (RH + 2) rows,
    // start (., -1, h0 - 1, n): the halo outside the feature map arrives as zeros) and the bytes land on `in_full`.  The box rows
    // are 128 bytes, SWIZZLE_128B: the 16-byte chunk j of position p sits at chunk j ^ (p & 7), so the 8 lanes of a quarter warp,
    // which read the same channel chunk of 8 consecutive positions, hit 8 different bank groups.
    const uint32_t tile_bytes = (uint32_t)(2 * hin_t * wp * CH * 4);
    auto fetch_tile = [&](int tile, float* xs, uint64_t* bar) {
      const int n = tile / tpu, h0 = (tile - n * tpu) * RH;
      mbar_expect_tx(bar, tile_bytes);
#pragma unroll
      for (int half = 0; half < 2; ++half)
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(smem_u32(xs + (size_t)half * XH)), "l"(&in_map), "r"(half * CH), "r"(-1), "r"(h0 - 1), "r"(n), "r"(smem_u32(bar))
                     : "memory");
    };
    if (tid == 0) {
      int t0 = blockIdx.x;
      if (t0 < ntiles) fetch_tile(t0, xs0, &in_full[0]);
      t0 += gridDim.x;
      if (t0 < ntiles) fetch_tile(t0, xs0 + (size_t)2 * XH, &in_full[1]);
    }
    // thread = (channel chunk, output column), the same for every tile (host side: C4 * wout <= producers): its nine taps and
    // the folded BatchNorm stay in registers
    const bool active = tid < C4 * L.wout;
    const int c4 = active ? tid / L.wout : 0, ow = active ? tid - c4 * L.wout : 0;
    const int half = c4 / CH4, cl = c4 - half * CH4;
    float4 kk[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) kk[j] = ld4(dws + j * C + 4 * c4);
    const float4 s = ld4(sc1 + 4 * c4), t = ld4(sf1 + 4 * c4);
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int b = it & 1, k = it >> 1;
      const int h0 = (tile % tpu) * RH, rh = imin(RH, L.hout - h0);
      float* xs = xs0 + (size_t)b * 2 * XH;
      float* a_hi = a0 + (size_t)b * 2 * C4 * LBA;
      float* a_lo = a_hi + (size_t)C4 * LBA;
      mbar_wait(&in_full[b], (uint32_t)(k & 1));
      if (it >= 2) mbar_wait(&mma_done[b], (uint32_t)((k - 1) & 1));             // the MMAs of tile it-2 have read this A buffer
      if (active) {
        const float* xh = xs + (size_t)half * XH;
        float* ah = a_hi + (size_t)c4 * LBA + 4 * ow;
        float* al = a_lo + (size_t)c4 * LBA + 4 * ow;
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0, acc2 = acc0;    // outputs r-2, r-1, r while input row r arrives
        int p = ow;                                               // position (tile row r, column ow) of the padded tile
        for (int r = 0; r < rh + 2; ++r, p += wp) {
          const float4 x0 = ld4(xh + p * CH + (((cl ^ p) & 7) << 2)), x1 = ld4(xh + (p + 1) * CH + (((cl ^ (p + 1)) & 7) << 2)),
                       x2 = ld4(xh + (p + 2) * CH + (((cl ^ (p + 2)) & 7) << 2));
#define TCR_DW3(acc, kr)                                                                                                                   \
  acc.x = fmaf(x0.x, kk[3 * (kr)].x, fmaf(x1.x, kk[3 * (kr) + 1].x, fmaf(x2.x, kk[3 * (kr) + 2].x, acc.x)));                               \
  acc.y = fmaf(x0.y, kk[3 * (kr)].y, fmaf(x1.y, kk[3 * (kr) + 1].y, fmaf(x2.y, kk[3 * (kr) + 2].y, acc.y)));                               \
  acc.z = fmaf(x0.z, kk[3 * (kr)].z, fmaf(x1.z, kk[3 * (kr) + 1].z, fmaf(x2.z, kk[3 * (kr) + 2].z, acc.z)));                               \
  acc.w = fmaf(x0.w, kk[3 * (kr)].w, fmaf(x1.w, kk[3 * (kr) + 1].w, fmaf(x2.w, kk[3 * (kr) + 2].w, acc.w)));
          TCR_DW3(acc0, 2)            // input row r is the bottom tap row of output r-2, the middle of r-1, the top of r
          TCR_DW3(acc1, 1)
          TCR_DW3(acc2, 0)
#undef TCR_DW3
          if (r >= 2) {
            float4 hi, lo;
            tf32_split(make_float4(fmaxf(fmaf(acc0.x, s.x, t.x), 0.f), fmaxf(fmaf(acc0.y, s.y, t.y), 0.f),
                                   fmaxf(fmaf(acc0.z, s.z, t.z), 0.f), fmaxf(fmaf(acc0.w, s.w, t.w), 0.f)), hi, lo);
            st4(ah + 4 * (r - 2) * L.wout, hi);
            st4(al + 4 * (r - 2) * L.wout, lo);
          }
          acc0 = acc1; acc1 = acc2; acc2 = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      fence_proxy_async();                                       // operand stores -> visible to the async proxy (UMMA)
      mbar_arrive(&a_full[b]);
      const int nxt = tile + 2 * (int)gridDim.x;                  // refill this input buffer once every producer has left it
      if (nxt < ntiles) {
        asm volatile("bar.sync 1, %0;" ::"n"(kWsProducers) : "memory");
        if (tid == 0) fetch_tile(nxt, xs, &in_full[b]);
      }
    }
  } else if (warp == kWsProducerWarps) {
    // ---- MMA issue: one thread ----
    if (lane == 0) ws_mma_role<CO>(ntiles, C4, LBA, a0, b_hi, b_lo, tmem, a_full, tmem_empty, mma_done);
  } else if (warp >= kTcThreads / 32 - 4) {
    // ---- epilogue: 4 warps ----
    ws_epilogue_role<CO>(L, RH, tpu, ntiles, tmem, sc2, sf2, out, mma_done, tmem_empty, warp & 3, lane);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
}

// First conv, warp-specialised (same roles and barriers as the block kernel): the producers copy the tile's input rows
// (cp.async, 4-byte cells, zero-filled halo), gather the im2col operand A[m][k] = x[oh*sh + k / kw][ow*sw + k % kw] (hi / lo) into
// the A buffer of the tile, and refill the input buffer for the tile two ahead; K = kh * kw (40).
template <int CO>
__global__ void __launch_bounds__(kTcThreads, 1) dscnn_conv_ws_kernel(DsLayerDev L, int RH, int n_utt, const float* __restrict__ params,
                                                                      const float* __restrict__ in, float* __restrict__ out, float eps) {
  TCR_DYNAMIC_SMEM(smem_raw);
  __shared__ uint64_t in_full[2], a_full[2], mma_done[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_s;
  float* smem = reinterpret_cast<float*>(smem_raw);
  constexpr int TMEM_COLS = 2 * CO <= 32 ? 32 : (2 * CO <= 64 ? 64 : (2 * CO <= 128 ? 128 : 256));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = L.kh * L.kw, K4 = K >> 2;
  const int hin_t = (RH - 1) * L.sh + L.kh, wp = (L.wout - 1) * L.sw + L.kw;
  const int XT = (hin_t * wp + 3) & ~3;
  constexpr int LBA = 128 * 4 + 4, LBB = CO * 4 + 4;
  float* a0 = smem;                                               // [2][hi, lo][K4][LBA]
  float* b_hi = a0 + (size_t)4 * K4 * LBA;                        // [K4][LBB]
  float* b_lo = b_hi + (size_t)K4 * LBB;
  float* sc = b_lo + (size_t)K4 * LBB;
  float* sf = sc + CO;
  float* xs0 = sf + CO;                                           // [2][hin_t][wp]
  const int tpu = (L.hout + RH - 1) / RH;
  const int ntiles = tpu * n_utt;
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&in_full[b], kWsProducers);
      mbar_init(&a_full[b], kWsProducers);
      mbar_init(&mma_done[b], 1);
      mbar_init(&tmem_empty[b], kWsFetchers);
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < K * CO; i += kTcThreads) {                // filters w[k][co] -> B operand [n = co][k]
    const int k = i / CO, co = i - k * CO;
    const float x = __ldg(params + L.w + i);
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    const int o = (k >> 2) * LBB + co * 4 + (k & 3);
    b_hi[o] = hi;
    b_lo[o] = x - hi;
  }
  for (int c = tid; c < CO; c += kTcThreads) fold_bn(params, L.b, L.beta, L.mm, L.mv, c, eps, sc, sf);
  pdl_wait();
  fence_proxy_async();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (warp < kWsProducerWarps) {
    // the cells a thread copies and the (position, tap chunk) operands it gathers are the same for every tile: worked out once
    constexpr int kCells = 4, kTasks = 4;                         // host side: hin_t * wp <= 4 * producers, 128 * K4 <= 4 * producers
    int f_pk[kCells];
#pragma unroll
    for (int j = 0; j < kCells; ++j) {
      const int i = tid + j * kWsProducers, r = i / wp, col = i - r * wp, x = col - L.pl;
      f_pk[j] = ((r * L.win + x + L.pl) << 10) | (r << 1) | ((i < hin_t * wp && x >= 0 && x < L.win) ? 1 : 0);
    }
    auto fetch_tile = [&](int tile, float* xs) {
      const int n = tile / tpu, h0 = (tile - n * tpu) * RH, hb = h0 * L.sh - L.pt;
      const float* base = in + ((int64_t)n * L.hin + hb) * L.win - L.pl;
#pragma unroll
      for (int j = 0; j < kCells; ++j) {
        if (tid + j * kWsProducers < hin_t * wp) {
          const int pk = f_pk[j], h = hb + ((pk >> 1) & 0x1FF);
          const bool ok = (pk & 1) && h >= 0 && h < L.hin;
          const uint32_t bytes = ok ? 4u : 0u;
          const float* src = ok ? base + (pk >> 10) : in;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(xs + tid + j * kWsProducers)), "l"(src), "r"(bytes) : "memory");
        }
      }
    };
    int g_x01[kTasks], g_x23[kTasks], g_a[kTasks];                // tile offsets of the four taps of a chunk (16 bits each), A offset | row
#pragma unroll
    for (int j = 0; j < kTasks; ++j) {
      const int task = tid + j * kWsProducers, m = task & 127, kc = task >> 7;
      const int oh = m / L.wout, ow = m - oh * L.wout;
      int e[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int kq = 4 * kc + q, ki = kq / L.kw, kj = kq - ki * L.kw;
        e[q] = (oh * L.sh + ki) * wp + ow * L.sw + kj;
      }
      g_x01[j] = e[0] | (e[1] << 16);
      g_x23[j] = e[2] | (e[3] << 16);
      g_a[j] = ((kc * LBA + 4 * m) << 7) | m;
    }
    {
      int t0 = blockIdx.x;
      if (t0 < ntiles) { fetch_tile(t0, xs0); cp_async_mbar_arrive(&in_full[0]); }
      t0 += gridDim.x;
      if (t0 < ntiles) { fetch_tile(t0, xs0 + XT); cp_async_mbar_arrive(&in_full[1]); }
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int b = it & 1, k = it >> 1;
      const int h0 = (tile % tpu) * RH, npos = imin(RH, L.hout - h0) * L.wout;
      float* xs = xs0 + b * XT;
      float* a_hi = a0 + (size_t)b * 2 * K4 * LBA;
      float* a_lo = a_hi + (size_t)K4 * LBA;
      mbar_wait(&in_full[b], (uint32_t)(k & 1));
      if (it >= 2) mbar_wait(&mma_done[b], (uint32_t)((k - 1) & 1));             // the MMAs of tile it-2 have read this A buffer
#pragma unroll
      for (int j = 0; j < kTasks; ++j) {
        if (tid + j * kWsProducers < 128 * K4) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if ((g_a[j] & 127) < npos) v = make_float4(xs[g_x01[j] & 0xFFFF], xs[g_x01[j] >> 16], xs[g_x23[j] & 0xFFFF], xs[g_x23[j] >> 16]);
          float4 hi, lo;
          tf32_split(v, hi, lo);
          st4(a_hi + (g_a[j] >> 7), hi);
          st4(a_lo + (g_a[j] >> 7), lo);
        }
      }
      fence_proxy_async();
      mbar_arrive(&a_full[b]);
      const int nxt = tile + 2 * (int)gridDim.x;
      if (nxt < ntiles) {
        asm volatile("bar.sync 1, %0;" ::"n"(kWsProducers) : "memory");
        fetch_tile(nxt, xs);
        cp_async_mbar_arrive(&in_full[b]);
      }
    }
  } else if (warp == kWsProducerWarps) {
    if (lane == 0) ws_mma_role<CO>(ntiles, K4, LBA, a0, b_hi, b_lo, tmem, a_full, tmem_empty, mma_done);
  } else if (warp >= kTcThreads / 32 - 4) {
    ws_epilogue_role<CO>(L, RH, tpu, ntiles, tmem, sc, sf, out, mma_done, tmem_empty, warp & 3, lane);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
}

// First layer (kh x kw conv from the single-channel feature map, audio_nets/ds_cnn.py:21 `conv_1`) as an implicit GEMM on
// tcgen05: A[m][k] = x[oh*sh + k / kw, ow*sw + k % kw] is gathered from the staged input rows straight into the canonical
// operand layout (hi / lo), B = the filter bank [co][k], K = kh * kw (40), same persistent / double-TMEM structure as above.
template <int CO>
__global__ void __launch_bounds__(kTcThreads, 1) dscnn_conv_tc_kernel(DsLayerDev L, int RH, int n_utt, const float* __restrict__ params,
                                                                      const float* __restrict__ in, float* __restrict__ out, float eps) {
  TCR_DYNAMIC_SMEM(smem_raw);
  __shared__ uint64_t mma_bar[2];
  __shared__ uint32_t tmem_base_s;
  float* smem = reinterpret_cast<float*>(smem_raw);
  constexpr int TMEM_COLS = 2 * CO <= 32 ? 32 : (2 * CO <= 64 ? 64 : (2 * CO <= 128 ? 128 : 256));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = L.kh * L.kw, K4 = K >> 2, KS = K >> 3;
  const int hin_t = (RH - 1) * L.sh + L.kh, wp = (L.wout - 1) * L.sw + L.kw;
  const int XT = (hin_t * wp + 3) & ~3;
  constexpr int LBA = 128 * 4 + 4, LBB = CO * 4 + 4;
  float* xs0 = smem;                                              // [2][hin_t][wp]
  float* a_hi = xs0 + 2 * XT;                                     // [K4][LBA]
  float* a_lo = a_hi + (size_t)K4 * LBA;
  float* b_hi = a_lo + (size_t)K4 * LBA;                          // [K4][LBB]
  float* b_lo = b_hi + (size_t)K4 * LBB;
  float* sc = b_lo + (size_t)K4 * LBB;
  float* sf = sc + CO;
  const int tpu = (L.hout + RH - 1) / RH;
  const int ntiles = tpu * n_utt;
  if (tid == 0) { mbar_init(&mma_bar[0], 1); mbar_init(&mma_bar[1], 1); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < K * CO; i += kTcThreads) {                // filters w[k][co] -> B operand [n = co][k]
    const int k = i / CO, co = i - k * CO;
    const float x = __ldg(params + L.w + i);
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    const int o = (k >> 2) * LBB + co * 4 + (k & 3);
    b_hi[o] = hi;
    b_lo[o] = x - hi;
  }
  for (int c = tid; c < CO; c += kTcThreads) fold_bn(params, L.b, L.beta, L.mm, L.mv, c, eps, sc, sf);
  pdl_wait();
  // The cells a thread copies and the (position, tap chunk) operands it gathers are the same for every tile: worked out once.
  constexpr int kCells = 4, kTasks = 4;                           // host side: hin_t * wp <= 4 * threads, 128 * K4 <= 4 * threads
  int f_pk[kCells];
#pragma unroll
  for (int j = 0; j < kCells; ++j) {
    const int i = tid + j * kTcThreads, r = i / wp, col = i - r * wp, x = col - L.pl;
    f_pk[j] = ((r * L.win + x + L.pl) << 10) | (r << 1) | ((i < hin_t * wp && x >= 0 && x < L.win) ? 1 : 0);
  }
  auto fetch_tile = [&](int tile, float* xs) {
    const int n = tile / tpu, h0 = (tile - n * tpu) * RH, hb = h0 * L.sh - L.pt;
    const float* base = in + ((int64_t)n * L.hin + hb) * L.win - L.pl;
#pragma unroll
    for (int j = 0; j < kCells; ++j) {
      if (tid + j * kTcThreads < hin_t * wp) {
        const int pk = f_pk[j], h = hb + ((pk >> 1) & 0x1FF);
        const bool ok = (pk & 1) && h >= 0 && h < L.hin;
        const uint32_t bytes = ok ? 4u : 0u;
        const float* src = ok ? base + (pk >> 10) : in;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(xs + tid + j * kTcThreads)), "l"(src), "r"(bytes) : "memory");
      }
    }
  };
  int g_x01[kTasks], g_x23[kTasks], g_a[kTasks];                  // tile offsets of the four taps of a chunk (16 bits each), A offset | row
#pragma unroll
  for (int j = 0; j < kTasks; ++j) {
    const int task = tid + j * kTcThreads, m = task & 127, kc = task >> 7;
    const int oh = m / L.wout, ow = m - oh * L.wout;
    int e[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = 4 * kc + q, ki = k / L.kw, kj = k - ki * L.kw;
      e[q] = (oh * L.sh + ki) * wp + ow * L.sw + kj;
    }
    g_x01[j] = e[0] | (e[1] << 16);
    g_x23[j] = e[2] | (e[3] << 16);
    g_a[j] = ((kc * LBA + 4 * m) << 7) | m;
  }
  auto epilogue = [&](uint32_t tacc, int tile_e) {
    const int n = tile_e / tpu, h0 = (tile_e - n * tpu) * RH, npos = imin(RH, L.hout - h0) * L.wout;
    constexpr int NG = kTcThreads / 128;
    const int m = 32 * (warp & 3) + lane, cbase = (warp >> 2) * (CO / NG);
    const size_t gbase = ((size_t)n * L.hout + h0) * L.wout;
#pragma unroll
    for (int c0 = 0; c0 < CO / NG; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tacc + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(cbase + c0);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (m < npos) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const int co = cbase + c0 + j;
          const float4 s = ld4(sc + co), t = ld4(sf + co);
          st4(out + (gbase + m) * CO + co,
              make_float4(fmaxf(fmaf(__uint_as_float(r[j]), s.x, t.x), 0.f), fmaxf(fmaf(__uint_as_float(r[j + 1]), s.y, t.y), 0.f),
                          fmaxf(fmaf(__uint_as_float(r[j + 2]), s.z, t.z), 0.f), fmaxf(fmaf(__uint_as_float(r[j + 3]), s.w, t.w), 0.f)));
        }
      }
    }
  };
  int tile = blockIdx.x, cur = 0, it = 0, prev_tile = -1;
  uint32_t parity[2] = {0u, 0u};
  if (tile < ntiles) fetch_tile(tile, xs0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CO >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t da_hi = umma_desc(smem_u32(a_hi), LBA * 4u, 128u), da_lo = umma_desc(smem_u32(a_lo), LBA * 4u, 128u);
  const uint64_t db_hi = umma_desc(smem_u32(b_hi), LBB * 4u, 128u), db_lo = umma_desc(smem_u32(b_lo), LBB * 4u, 128u);
  for (; tile < ntiles; tile += gridDim.x, cur ^= 1, ++it) {
    const int n = tile / tpu, h0 = (tile - n * tpu) * RH, rh = imin(RH, L.hout - h0);
    (void)n;
    const float* xs = xs0 + cur * XT;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (tile + (int)gridDim.x < ntiles) fetch_tile(tile + gridDim.x, xs0 + (cur ^ 1) * XT);
    if (prev_tile >= 0) {
      mbar_wait(&mma_bar[(it - 1) & 1], parity[(it - 1) & 1]);
      parity[(it - 1) & 1] ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // im2col gather: task = (position, 4-wide chunk of taps)
    const int npos = rh * L.wout;
#pragma unroll
    for (int j = 0; j < kTasks; ++j) {
      if (tid + j * kTcThreads < 128 * K4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((g_a[j] & 127) < npos) v = make_float4(xs[g_x01[j] & 0xFFFF], xs[g_x01[j] >> 16], xs[g_x23[j] & 0xFFFF], xs[g_x23[j] >> 16]);
        float4 hi, lo;
        tf32_split(v, hi, lo);
        st4(a_hi + (g_a[j] >> 7), hi);
        st4(a_lo + (g_a[j] >> 7), lo);
      }
    }
    fence_proxy_async();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      const uint32_t tacc = tmem + (uint32_t)((it & 1) * CO);
      for (int pass = 0; pass < 3; ++pass) {
        const uint64_t da0 = pass == 0 ? da_lo : da_hi, db0 = pass == 1 ? db_lo : db_hi;
        for (int ks = 0; ks < KS; ++ks) {
          const uint64_t da = da0 + (uint64_t)((2 * ks * LBA * 4) >> 4), db = db0 + (uint64_t)((2 * ks * LBB * 4) >> 4);
          const uint32_t accumulate = (pass | ks) ? 1u : 0u;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tacc), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_bar[it & 1])) : "memory");
    }
    if (prev_tile >= 0) epilogue(tmem + (uint32_t)(((it - 1) & 1) * CO), prev_tile);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    prev_tile = tile;
  }
  if (prev_tile >= 0) {
    mbar_wait(&mma_bar[(it - 1) & 1], parity[(it - 1) & 1]);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    epilogue(tmem + (uint32_t)(((it - 1) & 1) * CO), prev_tile);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
}
#endif

// ---- head: one CTA per utterance ----
__global__ void __launch_bounds__(256) dscnn_head_kernel(int npos, int C, int classes, int64_t fcw, int64_t fcb,
                                                         const float* __restrict__ params, const float* __restrict__ in,
                                                         float* __restrict__ logits, float* __restrict__ probs) {
  pdl_wait();
  __shared__ __align__(16) float s_red[1024];
  __shared__ float s_pool[320];
  __shared__ float s_logit[kMaxClasses];
  const int n = blockIdx.x, tid = threadIdx.x;
  int nseg;
  if ((C & 3) == 0) {                                             // float4 loads: thread = (segment of positions, 4 channels)
    const int C4 = C >> 2;
    nseg = imax(1, imin((int)blockDim.x / C4, 1024 / C));
    const int seg = tid / C4, c4 = tid - seg * C4;
    if (seg < nseg) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* src = in + (size_t)n * npos * C + 4 * c4;
#pragma unroll 4
      for (int p = seg; p < npos; p += nseg) {
        const float4 v = ldg4(src + (size_t)p * C);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      st4(s_red + seg * C + 4 * c4, s);
    }
  } else {
    nseg = imax(1, imin((int)blockDim.x / C, 1024 / C));
    const int seg = tid / C, c = tid - seg * C;
    if (seg < nseg) {
      float s = 0.f;
      for (int p = seg; p < npos; p += nseg) s += in[((size_t)n * npos + p) * C + c];
      s_red[seg * C + c] = s;
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) {
    float tot = 0.f;
    for (int q = 0; q < nseg; ++q) tot += s_red[q * C + c];
    s_pool[c] = tot / (float)npos;
  }
  __syncthreads();
  if (tid < classes) {
    float acc = params[fcb + tid];
    for (int k = 0; k < C; ++k) acc = fmaf(s_pool[k], params[fcw + (int64_t)k * classes + tid], acc);
    s_logit[tid] = acc;
    if (logits) logits[(size_t)n * classes + tid] = acc;
  }
  __syncthreads();
  if (tid < classes && probs) {
    float mx = -3.0e38f;
    for (int k = 0; k < classes; ++k) mx = fmaxf(mx, s_logit[k]);
    float se = 0.f;
    for (int k = 0; k < classes; ++k) se += expf(s_logit[k] - mx);
    probs[(size_t)n * classes + tid] = expf(s_logit[tid] - mx) / se;
  }
}

}  // namespace tcr

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace tcr;

struct tcr_dscnn {
  tcr_dscnn_config cfg;
  DsNet net;
  std::vector<tcr_param_desc> table;
  int64_t n_params = 0, flops = 0;
  float* act[2] = {nullptr, nullptr};
  size_t act_floats = 0;
  int sms = 148;
  bool use_tc = true;        // pointwise convs on tcgen05 where they fit (env TCR_DSCNN_TC=0: register-tiled FMA everywhere)
  bool tc_ws = true;         // warp-specialised tcgen05 block kernel (env TCR_DSCNN_TC=1: the lock-step version)
#ifndef TCR_EMU
  struct InMap { CUtensorMap map; const void* ptr = nullptr; int rh = 0; };
  std::vector<InMap> in_maps;   // per layer: TMA descriptor of the block's input activations (re-encoded if the buffer changes)
#endif
};

#ifndef TCR_EMU
// 4-D tensor map over the NHWC activations [n][h][w][c] (innermost first: c, w, h, n), box = 32 channels x wp columns x hin_t rows,
// 128-byte swizzle; out-of-range coordinates (the SAME-padding halo) are filled with zeros.  The driver entry point is looked up at
// run time (no link against libcuda).
static int ds_encode_in_map(CUtensorMap* map, const float* base, int c, int w, int h, int n, int box_w, int box_h) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled is not available");
      return TCR_ERR_CUDA;
    }
    encode = (EncodeFn)fn;
  }
  const cuuint64_t gdim[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  const cuuint64_t gstride[3] = {(cuuint64_t)c * 4, (cuuint64_t)w * c * 4, (cuuint64_t)h * w * c * 4};
  const cuuint32_t box[4] = {32u, (cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed"); return TCR_ERR_CUDA; }
  return TCR_OK;
}
#endif

static void ds_same(int len, int k, int s, int* out, int* lead) {
  *out = (len + s - 1) / s;
  const int total = std::max((*out - 1) * s + k - len, 0);
  *lead = total / 2;
}

extern "C" int tcr_dscnn_create(const tcr_dscnn_config* cfg, tcr_dscnn** out) {
  if (!cfg || !out) { set_error("NULL argument"); return TCR_ERR_INVALID; }
  *out = nullptr;
  if (cfg->size != 'S' && cfg->size != 'M' && cfg->size != 'L') { set_error("DS-CNN size must be 'S', 'M' or 'L'"); return TCR_ERR_INVALID; }
  if (cfg->max_batch <= 0 || cfg->height <= 0 || cfg->width <= 0 || cfg->num_classes <= 0 || cfg->num_classes > kMaxClasses) {
    set_error("bad DS-CNN configuration");
    return TCR_ERR_INVALID;
  }
  tcr_dscnn* d = new tcr_dscnn();
  d->cfg = *cfg;
  if (const char* e = getenv("TCR_DSCNN_TC")) { d->use_tc = e[0] != '0'; d->tc_ws = e[0] != '1'; }
  struct Def { int type, depth, kh, kw, sh, sw; const char* scope; };
  std::vector<Def> defs;
  if (cfg->size == 'S') {
    defs = {{0, 64, 10, 4, 2, 2, "conv_1"}, {1, 64, 3, 3, 1, 1, "conv_ds_1"}, {1, 64, 3, 3, 1, 1, "conv_ds_2"},
            {1, 64, 3, 3, 1, 1, "conv_ds_3"}, {1, 64, 3, 3, 1, 1, "conv_ds_4"}};
  } else if (cfg->size == 'M') {
    defs = {{0, 172, 10, 4, 2, 1, "conv_1"}, {1, 172, 3, 3, 2, 2, "conv_ds_1"}, {1, 172, 3, 3, 1, 1, "conv_ds_2"},
            {1, 172, 3, 3, 1, 1, "conv_ds_3"}, {1, 172, 3, 3, 1, 1, "conv_ds_4"}};
  } else {
    defs = {{0, 276, 10, 4, 2, 1, "conv_1"}, {1, 276, 3, 3, 2, 2, "conv_ds_1"}, {1, 276, 3, 3, 1, 1, "conv_ds_2"},
            {1, 276, 3, 3, 1, 1, "conv_ds_3"}, {1, 276, 3, 3, 1, 1, "conv_ds_4"}, {1, 276, 3, 3, 1, 1, "conv_ds_5"}};
  }
  int64_t off = 0;
  auto add = [&](const std::string& name, int kind, std::initializer_list<int> shape) {
    tcr_param_desc t;
    memset(&t, 0, sizeof(t));
    snprintf(t.name, sizeof(t.name), "%s", name.c_str());
    t.kind = kind;
    t.rank = (int)shape.size();
    int64_t numel = 1;
    int i = 0;
    for (int s : shape) { t.shape[i++] = s; numel *= s; }
    t.offset = off;
    t.numel = numel;
    d->table.push_back(t);
    off += numel;
    return t.offset;
  };
  int h = cfg->height, w = cfg->width, cin = 1;
  d->net.nlayers = 0;
  for (const Def& df : defs) {
    DsLayerDev L;
    memset(&L, 0, sizeof(L));
    L.type = df.type; L.cin = cin; L.cout = df.depth; L.kh = df.kh; L.kw = df.kw; L.sh = df.sh; L.sw = df.sw; L.hin = h; L.win = w;
    ds_same(h, df.kh, df.sh, &L.hout, &L.pt);
    ds_same(w, df.kw, df.sw, &L.wout, &L.pl);
    const std::string s = std::string("DSCNN/") + df.scope;
    if (df.type == 0) {
      L.w = add(s + "/weights", TCR_KIND_WEIGHT, {df.kh, df.kw, cin, df.depth});
      L.b = add(s + "/biases", TCR_KIND_BETA, {df.depth});
      L.beta = add(s + "/batch_norm/beta", TCR_KIND_BETA, {df.depth});
      L.mm = add(s + "/batch_norm/moving_mean", TCR_KIND_MOVING_MEAN, {df.depth});
      L.mv = add(s + "/batch_norm/moving_variance", TCR_KIND_MOVING_VAR, {df.depth});
      d->flops += 2ll * L.hout * L.wout * df.kh * df.kw * cin * df.depth;
    } else {
      L.w = add(s + "/depthwise_conv/depthwise_weights", TCR_KIND_WEIGHT, {df.kh, df.kw, cin, 1});
      L.b = add(s + "/depthwise_conv/biases", TCR_KIND_BETA, {cin});
      L.beta = add(s + "/dw_batch_norm/beta", TCR_KIND_BETA, {cin});
      L.mm = add(s + "/dw_batch_norm/moving_mean", TCR_KIND_MOVING_MEAN, {cin});
      L.mv = add(s + "/dw_batch_norm/moving_variance", TCR_KIND_MOVING_VAR, {cin});
      L.pw = add(s + "/pointwise_conv/weights", TCR_KIND_WEIGHT, {1, 1, cin, df.depth});
      L.pb = add(s + "/pointwise_conv/biases", TCR_KIND_BETA, {df.depth});
      L.pbeta = add(s + "/pw_batch_norm/beta", TCR_KIND_BETA, {df.depth});
      L.pmm = add(s + "/pw_batch_norm/moving_mean", TCR_KIND_MOVING_MEAN, {df.depth});
      L.pmv = add(s + "/pw_batch_norm/moving_variance", TCR_KIND_MOVING_VAR, {df.depth});
      d->flops += 2ll * L.hout * L.wout * df.kh * df.kw * cin + 2ll * L.hout * L.wout * cin * df.depth;
    }
    if (df.depth % 4) { delete d; set_error("DS-CNN depth must be a multiple of 4"); return TCR_ERR_UNSUPPORTED; }
    d->net.layer[d->net.nlayers++] = L;
    d->act_floats = std::max(d->act_floats, (size_t)L.hout * L.wout * df.depth);
    h = L.hout; w = L.wout; cin = df.depth;
  }
  d->net.classes = cfg->num_classes;
  d->net.fcw = add("DSCNN/fc1/weights", TCR_KIND_WEIGHT, {cin, cfg->num_classes});
  d->net.fcb = add("DSCNN/fc1/biases", TCR_KIND_BETA, {cfg->num_classes});
  d->flops += 2ll * cin * cfg->num_classes;
  d->n_params = off;
  if (cudaSetDevice(cfg->device) != cudaSuccess) { delete d; set_error("cudaSetDevice failed"); return TCR_ERR_CUDA; }
#ifndef TCR_EMU
  cudaDeviceGetAttribute(&d->sms, cudaDevAttrMultiProcessorCount, cfg->device);
#endif
  for (int i = 0; i < 2; ++i)
    if (cudaMalloc((void**)&d->act[i], d->act_floats * cfg->max_batch * sizeof(float)) != cudaSuccess) {
      tcr_dscnn_destroy(d);
      set_error("cudaMalloc failed for the DS-CNN activation buffers");
      return TCR_ERR_CUDA;
    }
  *out = d;
  return TCR_OK;
}

extern "C" int tcr_dscnn_destroy(tcr_dscnn* d) {
  if (!d) return TCR_OK;
  for (int i = 0; i < 2; ++i)
    if (d->act[i]) cudaFree(d->act[i]);
  delete d;
  return TCR_OK;
}

extern "C" int tcr_dscnn_param_table(const tcr_dscnn* d, const tcr_param_desc** descs, int32_t* count, int64_t* num_params,
                                     int64_t* forward_flops_per_utt) {
  if (!d || !descs || !count) { set_error("NULL argument"); return TCR_ERR_INVALID; }
  *descs = d->table.data();
  *count = (int32_t)d->table.size();
  if (num_params) *num_params = d->n_params;
  if (forward_flops_per_utt) *forward_flops_per_utt = d->flops;
  return TCR_OK;
}

extern "C" int tcr_dscnn_forward(tcr_dscnn* d, const float* features, const float* params, int32_t n, float* logits, float* probs,
                                 tcr_stream stream) {
  pdl_chain_reset();
  if (!d || !features || !params) { set_error("NULL argument"); return TCR_ERR_INVALID; }
  if (n <= 0 || n > d->cfg.max_batch) { set_error("n outside [1, max_batch]"); return TCR_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  const float eps = 1e-3f;
  const float* in = features;
  int cur = 0;
  for (int l = 0; l < d->net.nlayers; ++l) {
    const DsLayerDev& L = d->net.layer[l];
    float* out = d->act[cur];
    if (L.type == 0) {
#ifndef TCR_EMU
      int RHt = std::min(L.hout, 128 / std::max(1, std::min(L.wout, 128)));
      RHt = std::max(1, RHt);
      RHt = (L.hout + ((L.hout + RHt - 1) / RHt) - 1) / ((L.hout + RHt - 1) / RHt);
      const int K4 = L.kh * L.kw / 4, hin_t = (RHt - 1) * L.sh + L.kh, wpt = (L.wout - 1) * L.sw + L.kw;
      // implicit GEMM on tcgen05 (3xTF32); the limits are the kernel's precomputed per-thread walks (4 cells, 4 gather tasks)
      if (d->use_tc && L.cin == 1 && L.cout == 64 && (L.kh * L.kw) % 8 == 0 && L.wout <= 64 && hin_t * wpt <= 4 * kTcThreads &&
          128 * K4 <= 4 * kTcThreads && hin_t < 512) {
        const size_t smem_cws = (2 * (size_t)((hin_t * wpt + 3) & ~3) + 4 * (size_t)K4 * (128 * 4 + 4) + 2 * (size_t)K4 * (L.cout * 4 + 4) + 2 * L.cout) * 4;
        if (d->tc_ws && smem_cws <= 220 * 1024 && hin_t * wpt <= 4 * kWsProducers && 128 * K4 <= 4 * kWsProducers && (K4 & 1) == 0) {
          auto kws = dscnn_conv_ws_kernel<64>;
          static SmemOptIn optin_cws;
          if (optin_cws.ensure(kws, smem_cws) != cudaSuccess) return TCR_ERR_CUDA;
          const int tiles = ((L.hout + RHt - 1) / RHt) * n;
          TCR_LAUNCH("dscnn_conv_ws", kws, dim3(std::min(tiles, d->sms)), dim3(kTcThreads), smem_cws, s, L, RHt, n, params, in, out, eps);
          in = out;
          cur ^= 1;
          continue;
        }
        const size_t smem_tc = (2 * (size_t)((hin_t * wpt + 3) & ~3) + 2 * (size_t)K4 * (128 * 4 + 4) + 2 * (size_t)K4 * (L.cout * 4 + 4) + 2 * L.cout) * 4;
        auto ktc = dscnn_conv_tc_kernel<64>;
        static SmemOptIn optin_tc;
        if (optin_tc.ensure(ktc, smem_tc) != cudaSuccess) return TCR_ERR_CUDA;
        const int tiles = ((L.hout + RHt - 1) / RHt) * n;
        TCR_LAUNCH("dscnn_conv_tc", ktc, dim3(std::min(tiles, d->sms)), dim3(kTcThreads), smem_tc, s, L, RHt, n, params, in, out, eps);
        in = out;
        cur ^= 1;
        continue;
      }
#endif
      const int hp = (L.hout - 1) * L.sh + L.kh, wp = (L.wout - 1) * L.sw + L.kw;
      const size_t smem = (size_t)(((hp * wp + 3) & ~3) + L.kh * L.kw * L.cout + 2 * L.cout) * 4;
      auto kfn = dscnn_conv_kernel;
#ifndef TCR_EMU
      static SmemOptIn optin;
      if (optin.ensure(kfn, smem) != cudaSuccess) return TCR_ERR_CUDA;
#endif
      TCR_LAUNCH("dscnn_conv", kfn, dim3(n), dim3(256), smem, s, L, params, in, out, eps);
    } else {
      auto smem_for = [&](int RH, int COT) {
        const int hin_t = (RH - 1) * L.sh + L.kh, wp = (L.wout - 1) * L.sw + L.kw;
        return (size_t)((size_t)hin_t * wp * L.cin + (size_t)RH * L.wout * L.cin + (size_t)L.cin * COT + L.kh * L.kw * L.cin +
                        2 * L.cin + 2 * L.cout) * 4;
      };
#ifndef TCR_EMU
      // pointwise conv on the tensor cores (tcgen05, 3xTF32): the 64 -> 64 channel blocks with a 3x3 / stride-1 depthwise stage
      if (d->use_tc && L.cin == 64 && L.cout == 64 && L.kh == 3 && L.kw == 3 && L.sh == 1 && L.sw == 1 && L.wout <= 64) {
        int RHt = std::min(L.hout, 128 / L.wout);
        RHt = (L.hout + ((L.hout + RHt - 1) / RHt) - 1) / ((L.hout + RHt - 1) / RHt);     // balanced: 25 rows -> 5 tiles of 5
        const int C4 = L.cin / 4;
        const bool walk_ok = (RHt + 2) * (L.wout + 2) <= 8 * (kTcThreads / C4);      // the kernel's precomputed fetch walk: 8 cells
        const size_t smem_tc = (2 * (size_t)C4 * ((RHt + 2) * (L.wout + 2) * 4 + 4) + 2 * (size_t)C4 * (128 * 4 + 4) + 2 * (size_t)C4 * (L.cout * 4 + 4) +
                                (size_t)9 * L.cin + 2 * L.cin + 2 * L.cout) * 4;
        const int lba = ((RHt * L.wout + 7) & ~7) * 4 + 4;
        const size_t xh = (((size_t)(RHt + 2) * (L.wout + 2) * 128 + 1023) & ~(size_t)1023) / 4;       // floats per half input tile
        const size_t smem_ws = (4 * xh + 4 * (size_t)C4 * lba + 2 * (size_t)C4 * (L.cout * 4 + 4) + (size_t)9 * L.cin + 2 * L.cin + 2 * L.cout) * 4;
        if (d->tc_ws && smem_ws <= 226 * 1024 && C4 * L.wout <= kWsProducers && L.wout + 2 <= 256 && RHt + 2 <= 256 && in != features) {
          auto kws = dscnn_dsblock_ws_kernel<64, 64>;
          static SmemOptIn optin_ws;
          if (optin_ws.ensure(kws, smem_ws) != cudaSuccess) return TCR_ERR_CUDA;
          if ((int)d->in_maps.size() < d->net.nlayers) d->in_maps.resize(d->net.nlayers);
          tcr_dscnn::InMap& im = d->in_maps[l];
          if (im.ptr != in || im.rh != RHt) {                    // the activations ping-pong between two buffers: encoded once per layer
            const int rc = ds_encode_in_map(&im.map, in, L.cin, L.win, L.hin, d->cfg.max_batch, L.wout + 2, RHt + 2);
            if (rc != TCR_OK) return rc;
            im.ptr = in;
            im.rh = RHt;
          }
          const int tiles = ((L.hout + RHt - 1) / RHt) * n;
          TCR_LAUNCH("dscnn_dsblock_ws", kws, dim3(std::min(tiles, d->sms)), dim3(kTcThreads), smem_ws, s, im.map, L, RHt, n, params, out, eps);
          in = out;
          cur ^= 1;
          continue;
        }
        if (smem_tc <= 200 * 1024 && walk_ok) {
          auto ktc = dscnn_dsblock_tc_kernel<64, 64>;
          static SmemOptIn optin_tc;
          if (optin_tc.ensure(ktc, smem_tc) != cudaSuccess) return TCR_ERR_CUDA;
          const int tiles = ((L.hout + RHt - 1) / RHt) * n;
          TCR_LAUNCH("dscnn_dsblock_tc", ktc, dim3(std::min(tiles, d->sms)), dim3(kTcThreads), smem_tc, s, L, RHt, n, params, in, out, eps);
          in = out;
          cur ^= 1;
          continue;
        }
      }
#endif
      // the whole pointwise bank in shared memory when it fits next to a useful tile; else (DS-CNN-L: 276 x 276 fp32 = 305 KB) output-
      // channel tiles of COT, the depthwise output of the tile being reused for every slice
      int COT = L.cout;
      while (COT > 32 && smem_for(1, COT) > 96 * 1024) COT = ((COT / 2) + 3) & ~3;
      int RH = std::min(L.hout, 8);
      while (RH > 1 && smem_for(RH, COT) > 100 * 1024) --RH;
      const size_t smem = smem_for(RH, COT);
      if (smem > 200 * 1024) { set_error("DS-CNN layer does not fit in shared memory"); return TCR_ERR_UNSUPPORTED; }
      auto kfn = dscnn_dsblock_kernel;
#ifndef TCR_EMU
      static SmemOptIn optin;
      if (optin.ensure(kfn, smem) != cudaSuccess) return TCR_ERR_CUDA;
#endif
      TCR_LAUNCH("dscnn_dsblock", kfn, dim3((L.hout + RH - 1) / RH, n), dim3(256), smem, s, L, RH, COT, params, in, out, eps);
    }
    in = out;
    cur ^= 1;
  }
  const DsLayerDev& last = d->net.layer[d->net.nlayers - 1];
  if (last.cout > 320) { set_error("DS-CNN head supports up to 320 channels"); return TCR_ERR_UNSUPPORTED; }
  TCR_LAUNCH("dscnn_head", dscnn_head_kernel, dim3(n), dim3(256), 0, s, last.hout * last.wout, last.cout, d->net.classes, d->net.fcw,
             d->net.fcb, params, in, logits, probs);
  if (cudaGetLastError() != cudaSuccess) { set_error("DS-CNN launch failed"); return TCR_ERR_CUDA; }
  return TCR_OK;
}
