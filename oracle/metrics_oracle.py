"""TEST INFRASTRUCTURE (never imported by tc-resnet_b200/): NumPy restatement of the count-based evaluation metrics of the
reference, to check tcr_eval_accumulate (csrc/tcr_metrics.cu) and metrics.manager.metrics_from_counts.

Follows metrics/ops/non_tensor_ops.py: accuracy = sklearn accuracy_score(labels, predictions) (:64-101) with
labels / predictions = arg-max of the one-hot rows (metrics/parser.py:135-147), top-5 (:104-142, common/utils.py topN_accuracy:
the true class is among the N highest scores), precision / recall / F1 per class (:146-295), confusion matrix.
Parity unpinned: no reference vectors exist; the restatement is checked against direct definitions in tests/test_metrics.py."""
import numpy as np


def eval_counts(scores, onehot, topk=5):
    scores, onehot = np.asarray(scores), np.asarray(onehot)
    n, c = scores.shape
    y, p = onehot.argmax(1), scores.argmax(1)
    counts = np.zeros(c * c + 2, np.int64)
    np.add.at(counts, y * c + p, 1)
    sy = scores[np.arange(n), y]
    rank = (scores > sy[:, None]).sum(1) + ((scores == sy[:, None]) & (np.arange(c)[None, :] < y[:, None])).sum(1)
    counts[c * c] = int((rank < topk).sum())
    counts[c * c + 1] = n
    return counts
