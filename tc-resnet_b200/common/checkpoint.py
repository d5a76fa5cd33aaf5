"""Checkpoint I/O keyed by the reference's TF variable names.

The reference saves TF-V2 tensor bundles with tf.train.Saver every `step_save_checkpoint` steps as
`<train_dir>/<Model>-<step>` (helper/trainer.py:406-414) and the evaluator process discovers them through the
directory (common/tf_utils.py:65-67, 219-247).  The tensor-bundle format itself is a "next" row;
here the same names / directory protocol are kept with one `.npz` per checkpoint:
  <train_dir>/<Model>-<step>.npz   {TF variable name: array, "<var>/Momentum": array, "global_step": int64}
"""
from __future__ import annotations

import os
import re
import time
from pathlib import Path
from typing import Dict, Iterator, Optional

import numpy as np

_STEP = re.compile(r"-(\d+)\.npz$")


def checkpoint_step(path) -> int:
    """Global step parsed from the file name (common/tf_utils.py:237-247); 0 when absent."""
    m = _STEP.search(str(path))
    return int(m.group(1)) if m else 0


def save(train_dir, model_name: str, step: int, variables: Dict[str, np.ndarray], max_to_keep: int = 5, fmt: str = "npz") -> str:
    d = Path(train_dir)
    d.mkdir(parents=True, exist_ok=True)
    path = d / f"{model_name}-{int(step)}.npz"
    # unique per process and NOT matching "*.npz": neither another rank nor the watching evaluator can pick it up
    tmp = d / f".{model_name}-{int(step)}.{os.getpid()}.tmp"
    with open(tmp, "wb") as f:
        np.savez(f, global_step=np.int64(step), **variables)
    os.replace(tmp, path)                       # atomic: the watching evaluator never sees a partial file
    kept = sorted(d.glob(f"{model_name}-*.npz"), key=checkpoint_step)
    for old in kept[:-max_to_keep] if max_to_keep > 0 else []:
        old.unlink(missing_ok=True)
    return str(path)


def load(path) -> Dict[str, np.ndarray]:
    with np.load(str(path)) as z:
        return {k: z[k] for k in z.files}


def latest_checkpoint(directory) -> Optional[str]:
    cands = sorted((p for p in Path(directory).glob("*-*.npz") if not p.name.startswith(".") and _STEP.search(p.name)),
                   key=checkpoint_step)
    return str(cands[-1]) if cands else None


def resolve_checkpoint_path(path: str) -> str:
    """A directory resolves to its latest checkpoint (common/tf_utils.py:219-234)."""
    if path and Path(path).is_dir():
        return latest_checkpoint(path) or ""
    return path or ""


def checkpoints_iterator(directory, min_interval_secs: float = 0.0, timeout: Optional[float] = None) -> Iterator[str]:
    """Yields each new checkpoint as it appears (tf.contrib.training.checkpoints_iterator semantics)."""
    seen, waited = None, 0.0
    while True:
        newest = latest_checkpoint(directory)
        if newest is not None and newest != seen:
            seen, waited = newest, 0.0
            yield newest
            time.sleep(min_interval_secs)
            continue
        if timeout is not None and waited >= timeout:
            return
        time.sleep(1.0)
        waited += 1.0
