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
#include <stdio.h>
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
__global__ void __launch_bounds__(256) dscnn_dsblock_kernel(DsLayerDev L, int RH, const float* __restrict__ params,
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
  float* pws = ds + (size_t)RH * L.wout * C;                     // [C][CO]       pointwise weights
  float* dws = pws + (size_t)C * CO;                             // [kh*kw][C]    depthwise weights
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
  for (int i = tid; i < C * CO / 4; i += blockDim.x) st4(pws + 4 * i, ldg4(params + L.pw + 4 * i));
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
  // pointwise: tasks of kDsTM positions x 4 output channels, positions of a task are NRT apart
  const int ncg = CO >> 2, NRT = (npos + kDsTM - 1) / kDsTM;
  const size_t gbase = ((size_t)n * L.hout + h0) * L.wout;
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
      const float4 w0 = ld4(wk + (ci + 0) * CO), w1 = ld4(wk + (ci + 1) * CO), w2 = ld4(wk + (ci + 2) * CO), w3 = ld4(wk + (ci + 3) * CO);
#pragma unroll
      for (int i = 0; i < kDsTM; ++i) {
        const float4 x = ld4(xr[i] + ci);
        acc[i].x = fmaf(x.x, w0.x, acc[i].x); acc[i].y = fmaf(x.x, w0.y, acc[i].y); acc[i].z = fmaf(x.x, w0.z, acc[i].z); acc[i].w = fmaf(x.x, w0.w, acc[i].w);
        acc[i].x = fmaf(x.y, w1.x, acc[i].x); acc[i].y = fmaf(x.y, w1.y, acc[i].y); acc[i].z = fmaf(x.y, w1.z, acc[i].z); acc[i].w = fmaf(x.y, w1.w, acc[i].w);
        acc[i].x = fmaf(x.z, w2.x, acc[i].x); acc[i].y = fmaf(x.z, w2.y, acc[i].y); acc[i].z = fmaf(x.z, w2.z, acc[i].z); acc[i].w = fmaf(x.z, w2.w, acc[i].w);
        acc[i].x = fmaf(x.w, w3.x, acc[i].x); acc[i].y = fmaf(x.w, w3.y, acc[i].y); acc[i].z = fmaf(x.w, w3.z, acc[i].z); acc[i].w = fmaf(x.w, w3.w, acc[i].w);
      }
    }
    const float4 s = ld4(sc2 + 4 * cg), t = ld4(sf2 + 4 * cg);
#pragma unroll
    for (int i = 0; i < kDsTM; ++i) {
      const int pos = rt + i * NRT;
      if (pos < npos)
        st4(out + (gbase + pos) * CO + 4 * cg,
            make_float4(fmaxf(fmaf(acc[i].x, s.x, t.x), 0.f), fmaxf(fmaf(acc[i].y, s.y, t.y), 0.f),
                        fmaxf(fmaf(acc[i].z, s.z, t.z), 0.f), fmaxf(fmaf(acc[i].w, s.w, t.w), 0.f)));
    }
  }
}

// ---- head: one CTA per utterance ----
__global__ void __launch_bounds__(256) dscnn_head_kernel(int npos, int C, int classes, int64_t fcw, int64_t fcb,
                                                         const float* __restrict__ params, const float* __restrict__ in,
                                                         float* __restrict__ logits, float* __restrict__ probs) {
  pdl_wait();
  __shared__ float s_red[256];
  __shared__ float s_pool[320];
  __shared__ float s_logit[kMaxClasses];
  const int n = blockIdx.x, tid = threadIdx.x;
  const int nseg = imax(1, (int)blockDim.x / C), seg = tid / C, c = tid - seg * C;
  float s = 0.f;
  if (seg < nseg)
    for (int p = seg; p < npos; p += nseg) s += in[((size_t)n * npos + p) * C + c];
  if (seg < nseg) s_red[seg * C + c] = s;
  __syncthreads();
  if (tid < C) {
    float tot = 0.f;
    for (int q = 0; q < nseg; ++q) tot += s_red[q * C + tid];
    s_pool[tid] = tot / (float)npos;
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
};

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
      const int hp = (L.hout - 1) * L.sh + L.kh, wp = (L.wout - 1) * L.sw + L.kw;
      const size_t smem = (size_t)(((hp * wp + 3) & ~3) + L.kh * L.kw * L.cout + 2 * L.cout) * 4;
      auto kfn = dscnn_conv_kernel;
#ifndef TCR_EMU
      static SmemOptIn optin;
      if (optin.ensure(kfn, smem) != cudaSuccess) return TCR_ERR_CUDA;
#endif
      TCR_LAUNCH("dscnn_conv", kfn, dim3(n), dim3(256), smem, s, L, params, in, out, eps);
    } else {
      auto smem_for = [&](int RH) {
        const int hin_t = (RH - 1) * L.sh + L.kh, wp = (L.wout - 1) * L.sw + L.kw;
        return (size_t)((size_t)hin_t * wp * L.cin + (size_t)RH * L.wout * L.cin + (size_t)L.cin * L.cout + L.kh * L.kw * L.cin +
                        2 * L.cin + 2 * L.cout) * 4;
      };
      int RH = std::min(L.hout, 8);
      while (RH > 1 && smem_for(RH) > 100 * 1024) --RH;
      const size_t smem = smem_for(RH);
      if (smem > 200 * 1024) { set_error("DS-CNN layer does not fit in shared memory"); return TCR_ERR_UNSUPPORTED; }
      auto kfn = dscnn_dsblock_kernel;
#ifndef TCR_EMU
      static SmemOptIn optin;
      if (optin.ensure(kfn, smem) != cudaSuccess) return TCR_ERR_CUDA;
#endif
      TCR_LAUNCH("dscnn_dsblock", kfn, dim3((L.hout + RH - 1) / RH, n), dim3(256), smem, s, L, RH, params, in, out, eps);
    }
    in = out;
    cur ^= 1;
  }
  const DsLayerDev& last = d->net.layer[d->net.nlayers - 1];
  if (last.cout > 256) { set_error("DS-CNN head supports up to 256 channels"); return TCR_ERR_UNSUPPORTED; }
  TCR_LAUNCH("dscnn_head", dscnn_head_kernel, dim3(n), dim3(256), 0, s, last.hout * last.wout, last.cout, d->net.classes, d->net.fcw,
             d->net.fcb, params, in, logits, probs);
  if (cudaGetLastError() != cudaSuccess) { set_error("DS-CNN launch failed"); return TCR_ERR_CUDA; }
  return TCR_OK;
}
