"""Pure-Python mirror of the layer plan the C library builds (csrc/tcr_api.cu: build_plan), used by the
reference-facing shims and by bench.py for algorithmic FLOP / byte accounting.
Topology: audio_nets/tc_resnet.py:6-70 (channel plans :57-70, SAME padding, stride-2 blocks with a 1x1 shortcut)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional


@dataclass
class Conv:
    name: str
    cin: int
    cout: int
    k: int
    stride: int
    t_in: int
    t_out: int

    @property
    def weights(self) -> int:
        return self.k * self.cin * self.cout

    @property
    def macs(self) -> int:
        return self.t_out * self.k * self.cin * self.cout


@dataclass
class Block:
    down: Optional[Conv]
    conv_a: Conv
    conv_b: Conv


@dataclass
class Plan:
    scope: str
    frames: int
    features: int
    fft: int
    window: int
    stride: int
    clip: int
    conv0: Conv
    blocks: List[Block]
    c_last: int
    t_last: int
    num_classes: int

    def convs(self) -> List[Conv]:
        out = [self.conv0]
        for b in self.blocks:
            if b.down is not None:
                out.append(b.down)
            out += [b.conv_a, b.conv_b]
        return out

    @property
    def num_trainable(self) -> int:
        return sum(c.weights + 2 * c.cout for c in self.convs()) + self.c_last * (self.num_classes + 2)

    @property
    def forward_flops(self) -> int:
        return 2 * sum(c.macs for c in self.convs()) + 2 * self.c_last * (self.num_classes + 2)

    def frontend_flops(self, num_mel_bins=64, mel_nnz=None) -> float:
        """per frame 2.5 fft log2 fft + 3 bins + 2 nnz_mel + 2*64*F + W."""
        bins = self.fft // 2 + 1
        nnz = mel_nnz if mel_nnz is not None else {513: 942, 257: 471}.get(bins, 2 * bins)
        per_frame = 2.5 * self.fft * math.log2(self.fft) + 3 * bins + 2 * nnz + 2 * num_mel_bins * self.features + self.window
        return per_frame * self.frames

    def train_flops(self) -> float:
        """F_train = F_front + 3 F_fwd - F_conv0 (backward = 2x forward conv FLOPs, conv0 needs no dX)."""
        return self.frontend_flops() + 3 * self.forward_flops - 2 * self.conv0.macs

    def min_bytes(self, n_local: int) -> float:
        """B_min: compulsory HBM bytes per utterance: wav + label + logits + 24 B/param / N_local."""
        return 4 * self.clip + 4 * self.num_classes * 2 + 24.0 * self.num_trainable / n_local


def same_out(length: int, stride: int) -> int:
    return -(-length // stride)


def build_plan(model="TCResNet8", width_multiplier=1.0, sample_rate=16000, clip_duration_ms=1000, window_size_ms=40,
               window_stride_ms=20, features=40, num_classes=12) -> Plan:
    clip = int(sample_rate * clip_duration_ms / 1000)
    window = int(sample_rate * window_size_ms / 1000)
    stride = int(sample_rate * window_stride_ms / 1000)
    frames = 1 + (clip - window) // stride
    fft = 1 << (window - 1).bit_length()
    if model.startswith("TCResNet8"):
        scope, chans = "TCResNet8", [16, 24, 32, 48]
    elif model.startswith("TCResNet14"):
        scope, chans = "TCResNet14", [16, 24, 24, 32, 32, 48, 48]
    else:
        raise NotImplementedError(model)
    chans = [int(c * width_multiplier) for c in chans]
    conv0 = Conv("conv0", features, chans[0], 3, 1, frames, frames)
    blocks, c, t = [], chans[0], frames
    for i, n in enumerate(chans[1:]):
        s = 2 if n != c else 1
        down = Conv(f"block{i}/down", c, n, 1, 2, t, same_out(t, 2)) if n != c else None
        a = Conv(f"block{i}/conv{i}_0", c, n, 9, s, t, same_out(t, s))
        b = Conv(f"block{i}/conv{i}_1", n, n, 9, 1, a.t_out, a.t_out)
        blocks.append(Block(down, a, b))
        c, t = n, b.t_out
    return Plan(scope, frames, features, fft, window, stride, clip, conv0, blocks, c, t, num_classes)
