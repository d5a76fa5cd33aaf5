// tcr_bn.cuh — BatchNorm table helpers shared by forward and backward kernels.
// bnf layout: [4][C] = mean, rstd, scale (= gamma * rstd), beta.
#pragma once
#include "tcr_device.cuh"
#include "tcr_internal.h"

namespace tcr {

__device__ __forceinline__ float4 relu4(float4 v) {
  return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// z = (y - mean) * scale + beta
__device__ __forceinline__ float4 bn_apply4(float4 y, const float* __restrict__ bnf, int C, int c) {
  const float4 mean = ldc4(bnf + c), scale = ldc4(bnf + 2 * C + c), beta = ldc4(bnf + 3 * C + c);
  return make_float4(fmaf(y.x - mean.x, scale.x, beta.x), fmaf(y.y - mean.y, scale.y, beta.y),
                     fmaf(y.z - mean.z, scale.z, beta.z), fmaf(y.w - mean.w, scale.w, beta.w));
}
__device__ __forceinline__ float bn_apply1(float y, const float* __restrict__ bnf, int C, int c) {
  return fmaf(y - ldc1(bnf + c), ldc1(bnf + 2 * C + c), ldc1(bnf + 3 * C + c));
}
// xhat = (y - mean) * rstd
__device__ __forceinline__ float4 bn_xhat4(float4 y, const float* __restrict__ bnf, int C, int c) {
  const float4 mean = ldc4(bnf + c), rstd = ldc4(bnf + C + c);
  return make_float4((y.x - mean.x) * rstd.x, (y.y - mean.y) * rstd.y, (y.z - mean.z) * rstd.z, (y.w - mean.w) * rstd.w);
}

// Activation source: raw tensor or relu(bn(y)).
__device__ __forceinline__ float4 act_load4(const ActSrc& s, size_t ofs, int C, int c) {
  float4 v = ld4(s.data + ofs);
  if (s.kind == 1) v = relu4(bn_apply4(v, s.bnf, C, c));
  return v;
}

// dy = scale * (dz - s1/M - xhat * s2/M); dz optionally masked by the layer's own ReLU (bn(y) > 0).
__device__ __forceinline__ float4 dy_load4(const DySrc& d, size_t ofs, int C, int c) {
  float4 dz = ld4(d.dz + ofs);
  const float4 y = ld4(d.y + ofs);
  const float4 mean = ldc4(d.bnf + c), rstd = ldc4(d.bnf + C + c), scale = ldc4(d.bnf + 2 * C + c);
  if (d.mask_relu) {
    const float4 beta = ldc4(d.bnf + 3 * C + c);
    if (fmaf(y.x - mean.x, scale.x, beta.x) <= 0.f) dz.x = 0.f;
    if (fmaf(y.y - mean.y, scale.y, beta.y) <= 0.f) dz.y = 0.f;
    if (fmaf(y.z - mean.z, scale.z, beta.z) <= 0.f) dz.z = 0.f;
    if (fmaf(y.w - mean.w, scale.w, beta.w) <= 0.f) dz.w = 0.f;
  }
  const float4 s1 = ldc4(d.bsum + c), s2 = ldc4(d.bsum + C + c);
  const float im = d.inv_m;
  float4 r;
  r.x = scale.x * (dz.x - s1.x * im - (y.x - mean.x) * rstd.x * (s2.x * im));
  r.y = scale.y * (dz.y - s1.y * im - (y.y - mean.y) * rstd.y * (s2.y * im));
  r.z = scale.z * (dz.z - s1.z * im - (y.z - mean.z) * rstd.z * (s2.z * im));
  r.w = scale.w * (dz.w - s1.w * im - (y.w - mean.w) * rstd.w * (s2.w * im));
  return r;
}

// ---- per-channel constants hoisted into registers -------------------------------------------------
// Staging loops give every thread a FIXED group of 4 channels and walk rows, so the BN table entries (and the
// BN-backward sums) are loaded once per thread instead of once per element.
struct Chan4 { float4 mean, rstd, scale, beta; };
__device__ __forceinline__ Chan4 chan4_load(const float* __restrict__ bnf, int C, int c) {
  Chan4 k;
  k.mean = ldc4(bnf + c); k.rstd = ldc4(bnf + C + c); k.scale = ldc4(bnf + 2 * C + c); k.beta = ldc4(bnf + 3 * C + c);
  return k;
}
__device__ __forceinline__ float4 chan4_bn(const Chan4& k, float4 y) {
  return make_float4(fmaf(y.x - k.mean.x, k.scale.x, k.beta.x), fmaf(y.y - k.mean.y, k.scale.y, k.beta.y),
                     fmaf(y.z - k.mean.z, k.scale.z, k.beta.z), fmaf(y.w - k.mean.w, k.scale.w, k.beta.w));
}
__device__ __forceinline__ float4 chan4_xhat(const Chan4& k, float4 y) {
  return make_float4((y.x - k.mean.x) * k.rstd.x, (y.y - k.mean.y) * k.rstd.y, (y.z - k.mean.z) * k.rstd.z, (y.w - k.mean.w) * k.rstd.w);
}
// activation source with hoisted constants
struct Act4 { const float* data; Chan4 k; int kind; };
__device__ __forceinline__ Act4 act4_make(const ActSrc& s, int C, int c) {
  Act4 a;
  a.data = s.data; a.kind = s.kind;
  if (s.kind == 1) a.k = chan4_load(s.bnf, C, c);
  return a;
}
__device__ __forceinline__ float4 act4_load(const Act4& a, size_t ofs) {
  float4 v = ld4(a.data + ofs);
  if (a.kind == 1) v = relu4(chan4_bn(a.k, v));
  return v;
}
// dy = scale * (dz - s1/M - xhat * s2/M) with hoisted constants
struct Dy4 { const float* dz; const float* y; Chan4 k; float4 s1m, s2m; int mask; };
__device__ __forceinline__ Dy4 dy4_make(const DySrc& d, int C, int c) {
  Dy4 r;
  r.dz = d.dz; r.y = d.y; r.mask = d.mask_relu;
  r.k = chan4_load(d.bnf, C, c);
  const float4 s1 = ldc4(d.bsum + c), s2 = ldc4(d.bsum + C + c);
  const float im = d.inv_m;
  r.s1m = make_float4(s1.x * im, s1.y * im, s1.z * im, s1.w * im);
  r.s2m = make_float4(s2.x * im, s2.y * im, s2.z * im, s2.w * im);
  return r;
}
__device__ __forceinline__ float4 dy4_load(const Dy4& d, size_t ofs) {
  float4 dz = ld4(d.dz + ofs);
  const float4 y = ld4(d.y + ofs);
  if (d.mask) {
    const float4 z = chan4_bn(d.k, y);
    if (z.x <= 0.f) dz.x = 0.f;
    if (z.y <= 0.f) dz.y = 0.f;
    if (z.z <= 0.f) dz.z = 0.f;
    if (z.w <= 0.f) dz.w = 0.f;
  }
  float4 r;
  r.x = d.k.scale.x * (dz.x - d.s1m.x - (y.x - d.k.mean.x) * d.k.rstd.x * d.s2m.x);
  r.y = d.k.scale.y * (dz.y - d.s1m.y - (y.y - d.k.mean.y) * d.k.rstd.y * d.s2m.y);
  r.z = d.k.scale.z * (dz.z - d.s1m.z - (y.z - d.k.mean.z) * d.k.rstd.z * d.s2m.z);
  r.w = d.k.scale.w * (dz.w - d.s1m.w - (y.w - d.k.mean.w) * d.k.rstd.w * d.s2m.w);
  return r;
}

// 2-D walk of a [rows][c4n] float4 grid by the CTA: c4 fixed per thread, rows advance by rstep.
// Threads beyond rstep * c4n idle (row = huge).
struct RowWalk { int c4, row, rstep; };
__device__ __forceinline__ RowWalk row_walk(int tid, int nthreads, int c4n) {
  RowWalk w;
  w.rstep = nthreads / c4n;
  w.c4 = tid % c4n;
  w.row = tid / c4n;
  if (w.row >= w.rstep) w.row = 1 << 29;
  return w;
}

// ---- two-level "last CTA finalises" tree -------------------------------------------------------------
// A single CTA walking all G partials is a latency chain (measured: 22 us for G = 256).  Instead CTAs are
// grouped by kFanIn: the last CTA of each group to finish combines that group's <= 16 partials into a level-2
// record; the last group to finish combines the <= 16.. level-2 records into the table.  Every step reads at most
// kFanIn (or ceil(G/kFanIn)) records per channel with independent loads; grouping and order are fixed, so the
// result is deterministic.  counters: [0] = level-2 arrivals, [1 + g] = arrivals of group g (all self-resetting).
constexpr int kFanIn = 16;

// returns 0: nothing to do, 1: this CTA combines its group (level 1), then call tree_arrive_l2
__device__ __forceinline__ int tree_arrive_l1(unsigned* counters, int cta, int ncta) {
  __shared__ int s_flag;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int grp = cta / kFanIn;
    const int members = imin(kFanIn, ncta - grp * kFanIn);
    const unsigned prev = atomicAdd(counters + 1 + grp, 1u);
    s_flag = (prev == (unsigned)members - 1);
    if (s_flag) counters[1 + grp] = 0;
  }
  __syncthreads();
  const int f = s_flag;
  if (f) __threadfence();
  return f;
}
__device__ __forceinline__ int tree_arrive_l2(unsigned* counters, int ncta) {
  __shared__ int s_flag2;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ngrp = (ncta + kFanIn - 1) / kFanIn;
    const unsigned prev = atomicAdd(counters, 1u);
    s_flag2 = (prev == (unsigned)ngrp - 1);
    if (s_flag2) counters[0] = 0;
  }
  __syncthreads();
  const int f = s_flag2;
  if (f) __threadfence();
  return f;
}

// Level 1: sum the (sum y, sum y^2) partials of CTAs [g0, g1) into l2[(grp*C + c)*2 + q]; level 2: sum the l2 records and
// form the BN table.
__device__ __forceinline__ void bn_combine_l1(const BnFinalize& f, int grp, int ncta, int /*U*/, int /*n*/, int /*t_out*/, float* l2) {
  const int g0 = grp * kFanIn, g1 = imin(ncta, g0 + kFanIn);
  for (int i = threadIdx.x; i < 2 * f.c; i += blockDim.x) {
    float v[kFanIn];
#pragma unroll
    for (int j = 0; j < kFanIn; ++j) v[j] = __ldcg(f.fpart + (size_t)imin(g0 + j, g1 - 1) * f.c * 2 + i);
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < kFanIn; ++j)
      if (g0 + j < g1) s += (double)v[j];
    l2[(size_t)grp * f.c * 2 + i] = (float)s;
  }
}
__device__ __forceinline__ void bn_table_write(const BnFinalize& f, int c, double s1, double s2, double m_total, float eps) {
  const double mean = s1 / m_total;
  double var = s2 / m_total - mean * mean;
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + (double)eps);
  f.bnf[c] = (float)mean;
  f.bnf[f.c + c] = (float)rstd;
  f.bnf[2 * f.c + c] = (float)((double)f.gamma[c] * rstd);
  f.bnf[3 * f.c + c] = f.beta[c];
  f.var[c] = (float)var;
}
__device__ __forceinline__ void bn_combine_l2(const BnFinalize& f, int ngrp, const float* l2, float eps, double m_total) {
  for (int c = threadIdx.x; c < f.c; c += blockDim.x) {
    double s1 = 0.0, s2 = 0.0;
    for (int g0 = 0; g0 < ngrp; g0 += kFanIn) {           // all loads of a batch in flight before the first add
      float2 v[kFanIn];
#pragma unroll
      for (int j = 0; j < kFanIn; ++j) v[j] = __ldcg(reinterpret_cast<const float2*>(l2) + (size_t)imin(g0 + j, ngrp - 1) * f.c + c);
#pragma unroll
      for (int j = 0; j < kFanIn; ++j)
        if (g0 + j < ngrp) { s1 += (double)v[j].x; s2 += (double)v[j].y; }
    }
    bn_table_write(f, c, s1, s2, m_total, eps);
  }
}
// Sums of (sum dz, sum dz*xhat): level 1 -> l2[(grp*C + c)*2 + q], level 2 -> bsum[q*C + c]
__device__ __forceinline__ void bwdsum_combine_l1(const BwdSumFinalize& f, int grp, int ncta, float* l2) {
  const int g0 = grp * kFanIn, g1 = imin(ncta, g0 + kFanIn);
  for (int i = threadIdx.x; i < 2 * f.c; i += blockDim.x) {
    float v[kFanIn];
#pragma unroll
    for (int j = 0; j < kFanIn; ++j) v[j] = __ldcg(f.bpart + (size_t)imin(g0 + j, g1 - 1) * f.c * 2 + i);
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < kFanIn; ++j)
      if (g0 + j < g1) s += (double)v[j];
    l2[(size_t)grp * f.c * 2 + i] = (float)s;
  }
}
__device__ __forceinline__ void bwdsum_combine_l2(const BwdSumFinalize& f, int ngrp, const float* l2) {
  for (int i = threadIdx.x; i < 2 * f.c; i += blockDim.x) {
    double s = 0.0;
    for (int g0 = 0; g0 < ngrp; g0 += kFanIn) {
      float v[kFanIn];
#pragma unroll
      for (int j = 0; j < kFanIn; ++j) v[j] = __ldcg(l2 + (size_t)imin(g0 + j, ngrp - 1) * f.c * 2 + i);
#pragma unroll
      for (int j = 0; j < kFanIn; ++j)
        if (g0 + j < ngrp) s += (double)v[j];
    }
    f.bsum[(i & 1) * f.c + (i >> 1)] = (float)s;
  }
}

// Scalar (loss) variant of the tree: loss_part[G] -> l2[ngrp] -> *out
__device__ __forceinline__ void scalar_combine_l1(const float* part, int grp, int ncta, float* l2) {
  __shared__ float s_v[kFanIn];
  const int g0 = grp * kFanIn, g1 = imin(ncta, g0 + kFanIn);
  if (threadIdx.x < kFanIn) s_v[threadIdx.x] = (g0 + (int)threadIdx.x < g1) ? __ldcg(part + g0 + threadIdx.x) : 0.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int j = 0; j < kFanIn; ++j) s += (double)s_v[j];
    l2[grp] = (float)s;
  }
}
__device__ __forceinline__ void scalar_combine_l2(const float* l2, int ngrp, float* out) {
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int g = 0; g < ngrp; ++g) s += (double)__ldcg(l2 + g);
    *out = (float)s;
  }
}

// Per-channel (sum y, sum y^2) of a [rows][C] shared-memory tile -> part_out[c*2 + {0,1}], ONE pass.
// Requires C <= blockDim.x.  red: 2 * blockDim.x floats of scratch.  The cross-CTA combination is then a plain
// fixed-order sum; var = E[y^2] - mean^2 is formed once per channel in fp64 (relative error ~1e-7 (1 + mean^2/var)).
__device__ __forceinline__ void tile_stats(const float* tile, int rows, int C, float* red, float* /*unused*/, float* part_out) {
  const int tid = threadIdx.x;
  const int ns = imax(1, (int)blockDim.x / C);
  const int seg = tid / C, c = tid - seg * C;
  const bool act = seg < ns;
  if (act) {
    float s1 = 0.f, s2 = 0.f;
    for (int r = seg; r < rows; r += ns) {
      const float v = tile[r * C + c];
      s1 += v;
      s2 = fmaf(v, v, s2);
    }
    red[seg * C + c] = s1;
    red[blockDim.x + seg * C + c] = s2;
  }
  __syncthreads();
  if (tid < 2 * C) {
    const int q = tid / C, cc = tid - q * C;
    float tot = 0.f;
    for (int k = 0; k < ns; ++k) tot += red[q * blockDim.x + k * C + cc];
    part_out[(size_t)cc * 2 + q] = tot;
  }
  __syncthreads();
}

}  // namespace tcr
