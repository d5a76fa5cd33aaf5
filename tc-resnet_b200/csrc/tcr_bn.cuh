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
  const float4 mean = ldg4(bnf + c), scale = ldg4(bnf + 2 * C + c), beta = ldg4(bnf + 3 * C + c);
  return make_float4(fmaf(y.x - mean.x, scale.x, beta.x), fmaf(y.y - mean.y, scale.y, beta.y),
                     fmaf(y.z - mean.z, scale.z, beta.z), fmaf(y.w - mean.w, scale.w, beta.w));
}
__device__ __forceinline__ float bn_apply1(float y, const float* __restrict__ bnf, int C, int c) {
  return fmaf(y - __ldg(bnf + c), __ldg(bnf + 2 * C + c), __ldg(bnf + 3 * C + c));
}
// xhat = (y - mean) * rstd
__device__ __forceinline__ float4 bn_xhat4(float4 y, const float* __restrict__ bnf, int C, int c) {
  const float4 mean = ldg4(bnf + c), rstd = ldg4(bnf + C + c);
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
  const float4 mean = ldg4(d.bnf + c), rstd = ldg4(d.bnf + C + c), scale = ldg4(d.bnf + 2 * C + c);
  if (d.mask_relu) {
    const float4 beta = ldg4(d.bnf + 3 * C + c);
    if (fmaf(y.x - mean.x, scale.x, beta.x) <= 0.f) dz.x = 0.f;
    if (fmaf(y.y - mean.y, scale.y, beta.y) <= 0.f) dz.y = 0.f;
    if (fmaf(y.z - mean.z, scale.z, beta.z) <= 0.f) dz.z = 0.f;
    if (fmaf(y.w - mean.w, scale.w, beta.w) <= 0.f) dz.w = 0.f;
  }
  const float4 s1 = ldg4(d.bsum + c), s2 = ldg4(d.bsum + C + c);
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
  k.mean = ldg4(bnf + c); k.rstd = ldg4(bnf + C + c); k.scale = ldg4(bnf + 2 * C + c); k.beta = ldg4(bnf + 3 * C + c);
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
  const float4 s1 = ldg4(d.bsum + c), s2 = ldg4(d.bsum + C + c);
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

// "Last CTA finalises": returns true in exactly one CTA of the grid, after every other CTA's global
// writes are visible.  The counter is reset for the next launch.
__device__ __forceinline__ bool last_block_done(unsigned* counter, unsigned nblocks) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(counter, 1u);
    s_last = (prev == nblocks - 1);
    if (s_last) *counter = 0;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// Chan et al. combination of per-group (mean, M2) partials into the BN table (fp64, fixed order), spread over
// the whole CTA: thread (seg, c) owns every ns-th group of channel c, so each thread issues a handful of
// independent loads instead of one thread walking all groups (that serial chain was a 20-30 us kernel tail).
// Requires C <= blockDim.x <= 512.
__device__ __forceinline__ void bn_finalize(const BnFinalize& f, int groups, int U, int n, int t_out, float eps) {
  __shared__ double s_part[512];
  __shared__ double s_mean[128];
  const int tid = threadIdx.x, C = f.c;
  const int ns = imax(1, (int)blockDim.x / C);
  const int seg = tid / C, c = tid - seg * C;
  const bool act = seg < ns;
  const double m_total = (double)n * t_out;
  double s = 0.0;
  if (act)
    for (int g = seg; g < groups; g += ns)
      s += (double)(imin(U, n - g * U) * t_out) * (double)__ldcg(f.fpart + ((size_t)g * C + c) * 2);
  if (act) s_part[seg * C + c] = s;
  __syncthreads();
  if (tid < C) {
    double tot = 0.0;
    for (int q = 0; q < ns; ++q) tot += s_part[q * C + tid];
    s_mean[tid] = tot / m_total;
  }
  __syncthreads();
  double m2 = 0.0;
  if (act) {
    const double mean = s_mean[c];
    for (int g = seg; g < groups; g += ns) {
      const double cnt = (double)(imin(U, n - g * U) * t_out);
      const double d = (double)__ldcg(f.fpart + ((size_t)g * C + c) * 2) - mean;
      m2 += (double)__ldcg(f.fpart + ((size_t)g * C + c) * 2 + 1) + cnt * d * d;
    }
    s_part[seg * C + c] = m2;
  }
  __syncthreads();
  if (tid < C) {
    double tot = 0.0;
    for (int q = 0; q < ns; ++q) tot += s_part[q * C + tid];
    const double var = tot / m_total;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    f.bnf[tid] = (float)s_mean[tid];
    f.bnf[C + tid] = (float)rstd;
    f.bnf[2 * C + tid] = (float)((double)f.gamma[tid] * rstd);
    f.bnf[3 * C + tid] = f.beta[tid];
    f.var[tid] = (float)var;
  }
  __syncthreads();
}

// Sum of per-group (sum dz, sum dz*xhat) partials, same thread layout (2C columns).
__device__ __forceinline__ void bwdsum_finalize(const BwdSumFinalize& f, int groups) {
  __shared__ double s_part[512];
  const int tid = threadIdx.x, C2 = 2 * f.c;
  const int ns = imax(1, (int)blockDim.x / C2);
  const int seg = tid / C2, i = tid - seg * C2;
  const bool act = seg < ns;
  if (act) {
    const int c = i >> 1, q = i & 1;
    double s = 0.0;
    for (int g = seg; g < groups; g += ns) s += (double)__ldcg(f.bpart + ((size_t)g * f.c + c) * 2 + q);
    s_part[seg * C2 + i] = s;
  }
  __syncthreads();
  if (tid < C2) {
    double tot = 0.0;
    for (int q = 0; q < ns; ++q) tot += s_part[q * C2 + tid];
    f.bsum[(tid & 1) * f.c + (tid >> 1)] = (float)tot;
  }
  __syncthreads();
}

// Per-channel (mean, M2) of a [rows][C] shared-memory tile -> part_out[c*2 + {0,1}].
// Requires C <= blockDim.x.  red: blockDim.x floats of scratch, smean: C floats.  Two passes over the
// tile (mean, then squared deviations) so the partial is exact enough for the Chan combination.
__device__ __forceinline__ void tile_stats(const float* tile, int rows, int C, float* red, float* smean, float* part_out) {
  const int tid = threadIdx.x;
  const int ns = imax(1, (int)blockDim.x / C);
  const int seg = tid / C, c = tid - seg * C;
  const bool act = seg < ns;
  float s = 0.f;
  if (act) {
    for (int r = seg; r < rows; r += ns) s += tile[r * C + c];
    red[seg * C + c] = s;
  }
  __syncthreads();
  if (tid < C) {
    float tot = 0.f;
    for (int q = 0; q < ns; ++q) tot += red[q * C + tid];
    smean[tid] = tot / (float)rows;
  }
  __syncthreads();
  if (act) {
    const float mu = smean[c];
    float m2 = 0.f;
    for (int r = seg; r < rows; r += ns) {
      const float d = tile[r * C + c] - mu;
      m2 = fmaf(d, d, m2);
    }
    red[seg * C + c] = m2;
  }
  __syncthreads();
  if (tid < C) {
    float tot = 0.f;
    for (int q = 0; q < ns; ++q) tot += red[q * C + tid];
    part_out[(size_t)tid * 2] = smean[tid];
    part_out[(size_t)tid * 2 + 1] = tot;
  }
  __syncthreads();
}

}  // namespace tcr
