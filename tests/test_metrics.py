"""Evaluation metrics on the device: tcr_eval_accumulate vs the NumPy oracle (bit-exact: integer counts),
and metrics_from_counts vs the host metrics of the [num_samples, classes] path."""
import numpy as np
import pytest

import tcresnet_b200  # noqa: F401
from oracle import metrics_oracle as MO
from tcresnet_b200.metrics.manager import AudioMetricManager, metrics_from_counts, topn_accuracy
from tcr_harness import Engine, NumpyBackend


def _batch(rng, n, c=12, ties=True):
    scores = rng.standard_normal((n, c)).astype(np.float32)
    if ties:                                      # exact ties: arg-max takes the first maximum, ranks break ties by class index
        scores[::7, 3] = scores[::7, 5]
        scores[::11] = 0.25
    labels = rng.integers(0, c, n)
    return scores, np.eye(c, dtype=np.float32)[labels]


def _check(backend):
    rng = np.random.default_rng(3)
    eng = Engine(backend, max_batch=8)
    total = None
    ref = np.zeros(12 * 12 + 2, np.int64)
    for n in (1, 39, 300):                        # the reference's evaluation batch sizes are small and odd (3, 39)
        scores, onehot = _batch(rng, n)
        total = eng.eval_accumulate(scores, onehot, 5, total)
        ref += MO.eval_counts(scores, onehot, 5)
    assert np.array_equal(total, ref)             # integers: bit-exact
    assert total[-1] == 340 and total[:144].sum() == 340
    eng.close()


def test_eval_accumulate_matches_oracle_emulated():
    _check(NumpyBackend())


@pytest.mark.gpu
def test_eval_accumulate_matches_oracle_gpu():
    from tcr_harness import TorchBackend
    _check(TorchBackend())


def test_metrics_from_counts_equal_host_metrics():
    rng = np.random.default_rng(5)
    scores, onehot = _batch(rng, 500, ties=False)
    names = [f"c{i}" for i in range(12)]
    counts = MO.eval_counts(scores, onehot, 5)
    dev = metrics_from_counts(counts, names, use_class_metrics=True)
    host = AudioMetricManager(False, True, [])
    host.evaluate_and_aggregate_metrics(1, {"labels_onehot": onehot, "predictions_onehot": scores, "label_names": names}, {})
    h = host.get_evaluation_result(1)
    for k, v in dev.items():
        if k == "confusion_matrix":
            assert np.array_equal(v, h[k])
        else:
            assert abs(v - h[k]) < 1e-12, k
    assert abs(dev["top5_accuracy"] - topn_accuracy(onehot, scores, 5)) < 1e-12
