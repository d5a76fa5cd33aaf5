// tcr_metrics.cu — the evaluation consumer right behind the forward pass, on the device.
//
// The reference's evaluator pulls `predictions_onehot` [num_samples, classes] and the labels to the host after every batch
// (helper/base.py:52-143, metrics/parser.py:135-147) and computes accuracy / top-5 / precision / recall / F1 / the classification
// report with sklearn (metrics/ops/non_tensor_ops.py:64-142, :146-295, :346-).  All of these are functions of the confusion
// matrix plus one top-k counter, so the batch is reduced to classes^2 + 2 integers where the scores already are:
//   counts[y * C + p] += 1     y = arg-max of the one-hot label row, p = arg-max of the score row (first maximum, like np.argmax)
//   counts[C * C]     += 1     when the true class ranks among the top k scores (ties broken by the lower class index)
//   counts[C * C + 1] += 1     per utterance
// Integer atomics: exact and order-independent, so the result is bit-identical to the host computation.
#include "tcr_device.cuh"
#include "tcr_net.h"

namespace tcr {

__global__ void __launch_bounds__(256) eval_accumulate_kernel(const float* __restrict__ scores, const float* __restrict__ onehot, int n, int C,
                                                              int topk, unsigned long long* counts) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* s = scores + (size_t)i * C;
  const float* l = onehot + (size_t)i * C;
  int y = 0, p = 0;
  float ly = l[0], sp = s[0];
  for (int c = 1; c < C; ++c) {
    if (l[c] > ly) { ly = l[c]; y = c; }
    if (s[c] > sp) { sp = s[c]; p = c; }
  }
  const float sy = s[y];
  int rank = 0;                       // classes ranked before the true one by a stable descending sort
  for (int c = 0; c < C; ++c) rank += (s[c] > sy || (s[c] == sy && c < y)) ? 1 : 0;
  atomicAdd(counts + (size_t)y * C + p, 1ull);
  if (rank < topk) atomicAdd(counts + (size_t)C * C, 1ull);
  atomicAdd(counts + (size_t)C * C + 1, 1ull);
}

int eval_accumulate_launch(const float* scores, const float* onehot, int n, int classes, int topk, int64_t* counts, cudaStream_t s) {
  TCR_LAUNCH("eval_accumulate", eval_accumulate_kernel, dim3((n + 255) / 256), dim3(256), 0, s, scores, onehot, n, classes, topk,
             reinterpret_cast<unsigned long long*>(counts));
  return 0;
}

}  // namespace tcr
