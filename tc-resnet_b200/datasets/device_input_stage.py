"""The per-clip input stage on the device (SURVEY.md 8f "next" row 1): host side.

The reference runs decode_wav, pad/crop, time shift, background mix and clipping per file on `tf.data` host threads
(datasets/augmentation_factory.py:30-211, datasets/audio_data_wrapper.py:37-58).  Here the host only makes the random DRAWS
(24 bytes per clip) and ships the wav files' own int16 samples; tcr_augment_pcm16 (csrc/tcr_augment.cu) does the arithmetic for the
whole batch in one HBM-bound pass, bit-identical to the host stage (augmentation_factory.py in this package) for the same draws.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

# tcr_augment_clip (include/tcr_b200.h)
CLIP_DTYPE = np.dtype([("length", "<i4"), ("shift", "<i4"), ("silent", "<i4"), ("bg_volume", "<f4"), ("bg_offset", "<i8")])


def draw_clips(rng: np.random.RandomState, lengths: Sequence[int], silent: Sequence[bool], clip_samples: int,
               bg_lengths: Sequence[int], shift: bool = True, is_training: bool = True, background_frequency: float = 0.8,
               background_max_volume: float = 0.1, shift_ratio: float = 0.1) -> np.ndarray:
    """One record per clip, drawn in the order of the reference's graph: time shift (augmentation_factory.py:104-112),
    background choice and random crop (:58-67), background volume ("naive" version, :69-80)."""
    n = len(lengths)
    clips = np.zeros(n, CLIP_DTYPE)
    starts = np.concatenate([[0], np.cumsum(bg_lengths)])[:-1] if len(bg_lengths) else np.zeros(0, np.int64)
    limit = int(clip_samples * shift_ratio)
    for i in range(n):
        clips[i]["length"] = int(lengths[i])
        clips[i]["silent"] = int(bool(silent[i]))
        clips[i]["shift"] = int(rng.randint(-limit, limit)) if (shift and limit) else 0
        clips[i]["bg_offset"] = -1
        if len(bg_lengths):
            b = int(rng.randint(0, len(bg_lengths)))
            # recordings shorter than a clip are zero-padded to clip length by DeviceInputStage (tf.random_crop would raise)
            clips[i]["bg_offset"] = int(starts[b]) + int(rng.randint(0, max(bg_lengths[b] - clip_samples, 0) + 1))
            mixed = is_training and rng.uniform() < background_frequency
            clips[i]["bg_volume"] = np.float32(rng.uniform(0.0, background_max_volume)) if mixed else np.float32(0.0)
    return clips


def pack(clips: np.ndarray, device) -> torch.Tensor:
    """The records as the uint8 CUDA tensor Engine.augment takes."""
    assert clips.dtype == CLIP_DTYPE
    return torch.from_numpy(np.frombuffer(np.ascontiguousarray(clips).tobytes(), np.uint8).copy()).to(device)


class DeviceInputStage:
    """Background bank resident in HBM + one call per batch: int16 clips and draws in, fp32 wav [n, clip_samples] out."""

    def __init__(self, engine, background_data: Optional[Sequence[np.ndarray]] = None):
        self.engine = engine
        clip = int(engine.cfg.clip_samples)
        bank = [np.asarray(b, np.float32).reshape(-1) for b in (background_data or [])]
        bank = [np.pad(b, (0, clip - b.shape[0])) if b.shape[0] < clip else b for b in bank]     # every crop fits its recording
        self.bg_lengths = [int(b.shape[0]) for b in bank]
        self.background = torch.from_numpy(np.concatenate(bank)).to(engine.device) if bank else None
        if bank:                                     # the kernel refuses crops that would run past the bank
            engine.lib.tcr_set_background_samples(engine._h, int(sum(self.bg_lengths)))

    def __call__(self, pcm: torch.Tensor, clips: np.ndarray, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.engine.augment(pcm, pack(clips, self.engine.device), self.background, out=out)
