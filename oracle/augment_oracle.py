"""CPU restatement (NumPy fp32) of the reference's per-clip input stage — TEST INFRASTRUCTURE ONLY.

Checker for the device input stage (tc-resnet_b200/csrc/tcr_augment.cu, tcr_augment_pcm16).  PARITY UNPINNED: the reference
holds no golden vectors for this stage and its random draws come from TF's RNG; the draws are therefore inputs here.
Follows datasets/augmentation_factory.py: decode + crop / zero-pad (:146-158), silent clips (:172-178), _shift_audio
(:104-143), _mix_background "naive" version (:30-101: tf.multiply, tf.add, tf.clip_by_value, all fp32).
"""
from __future__ import annotations

import numpy as np

CLIP_DTYPE = np.dtype([("length", "<i4"), ("shift", "<i4"), ("silent", "<i4"), ("bg_volume", "<f4"), ("bg_offset", "<i8")])   # tcr_augment_clip


def augment(pcm: np.ndarray, clips: np.ndarray, background: np.ndarray | None, clip_samples: int) -> np.ndarray:
    """pcm int16 [n, stride]; clips: CLIP_DTYPE [n]; background: concatenated fp32 recordings or None -> fp32 [n, clip_samples]."""
    n = pcm.shape[0]
    out = np.zeros((n, clip_samples), np.float32)
    for i in range(n):
        c = clips[i]
        x = np.zeros(clip_samples, np.float32)
        if not c["silent"]:
            m = min(int(c["length"]), clip_samples, pcm.shape[1])
            x[:m] = pcm[i, :m].astype(np.float32) / np.float32(32768.0)            # decode_wav, crop / zero-pad
        s = int(c["shift"])
        y = np.zeros_like(x)                                                        # _shift_audio: pad one side, slice the other
        if s >= 0:
            y[s:] = x[:clip_samples - s]
        else:
            y[:clip_samples + s] = x[-s:]
        if background is not None and c["bg_offset"] >= 0:
            o = int(c["bg_offset"])
            y = background[o:o + clip_samples].astype(np.float32) * np.float32(c["bg_volume"]) + y      # multiply, then add
        out[i] = np.clip(y, np.float32(-1.0), np.float32(1.0))
    return out


def random_clips(rng: np.random.RandomState, n: int, clip_samples: int, stride: int, bg_lengths, training=True,
                 background_frequency=0.8, background_max_volume=0.1, shift_ratio=0.1, silent_fraction=0.1):
    """Host-side random draws in the order the input stage makes them (shift, background choice, crop offset, volume)."""
    clips = np.zeros(n, CLIP_DTYPE)
    starts = np.concatenate([[0], np.cumsum(bg_lengths)])[:-1] if len(bg_lengths) else np.zeros(0, np.int64)
    limit = int(clip_samples * shift_ratio)
    for i in range(n):
        clips[i]["silent"] = int(rng.uniform() < silent_fraction)
        clips[i]["length"] = int(rng.randint(clip_samples // 2, stride + 1))
        clips[i]["shift"] = int(rng.randint(-limit, limit)) if limit else 0
        if len(bg_lengths):
            b = int(rng.randint(0, len(bg_lengths)))
            clips[i]["bg_offset"] = int(starts[b]) + int(rng.randint(0, bg_lengths[b] - clip_samples + 1))
            mixed = training and rng.uniform() < background_frequency
            clips[i]["bg_volume"] = np.float32(rng.uniform(0.0, background_max_volume)) if mixed else np.float32(0.0)
        else:
            clips[i]["bg_offset"] = -1
    return clips
