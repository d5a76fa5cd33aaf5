"""PyTorch-CPU fp32 port of the oracle (TEST / BENCH INFRASTRUCTURE ONLY — never imported by tc-resnet_b200/).

"restated reference, PyTorch-CPU (TF 1.13.1 unavailable)": the same arithmetic as oracle/tcr_oracle.py
(which cites the reference file:line per function), written with torch ops + autograd so that it uses all
host cores (oneDNN convs, MKL FFT).  bench.py times it as the CPU baseline / `--impl reference` arm; it is
checked against the NumPy oracle in tests/test_oracle.py.  PARITY UNPINNED (no reference goldens exist).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import tcr_oracle as O


class TorchPort:
    def __init__(self, spec: O.NetSpec, params, moving, window: int, stride: int, keep_prob: float = 0.5,
                 dtype=torch.float32, seed: int = 0):
        self.spec, self.window, self.stride, self.keep = spec, window, stride, keep_prob
        self.dtype = dtype
        npd = np.float64 if dtype == torch.float64 else np.float32
        self.p = {k: torch.tensor(np.asarray(v, npd), requires_grad=True) for k, v in params.items()}
        self.mv = {k: torch.tensor(np.asarray(v, npd)) for k, v in moving.items()}
        self.slots = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.fft = O.next_pow2(window)
        self.hann = torch.tensor(O.hann_window_periodic(window, npd))
        self.mel = torch.tensor(O.linear_to_mel_weight_matrix(64, self.fft // 2 + 1, dtype=npd))
        self.dct = torch.tensor(O.dct2_matrix(64, spec.f_in, npd))
        self.gen = torch.Generator().manual_seed(seed)

    def mfcc(self, wav: torch.Tensor) -> torch.Tensor:
        frames = wav.unfold(1, self.window, self.stride) * self.hann          # no centring / reflect padding
        spec = torch.fft.rfft(frames, n=self.fft, dim=-1)
        power = spec.real * spec.real + spec.imag * spec.imag
        return torch.log(power @ self.mel + 1e-6) @ self.dct

    def _conv_bn(self, x, cv, training, stats):
        p = f"{self.spec.scope}/{cv.name}"
        w = self.p[p + "/weights"][:, 0].permute(2, 1, 0)
        y = F.conv1d(F.pad(x, (cv.pad_left, cv.pad_right)), w, stride=cv.stride)
        if training:
            if stats is not None:
                with torch.no_grad():
                    m = y.shape[0] * y.shape[2]
                    stats[p] = (y.mean(dim=(0, 2)), y.var(dim=(0, 2), unbiased=False) * (m / max(m - 1, 1)))
            z = F.batch_norm(y, None, None, self.p[p + "/BatchNorm/gamma"], self.p[p + "/BatchNorm/beta"], True, 0.0, O.BN_EPS)
        else:
            z = F.batch_norm(y, self.mv[p + "/BatchNorm/moving_mean"], self.mv[p + "/BatchNorm/moving_variance"],
                             self.p[p + "/BatchNorm/gamma"], self.p[p + "/BatchNorm/beta"], False, 0.0, O.BN_EPS)
        return F.relu(z) if cv.relu else z

    def logits(self, feat: torch.Tensor, training: bool, stats=None, mask=None):
        net = self._conv_bn(feat.permute(0, 2, 1), self.spec.conv0, training, stats)
        for b in self.spec.blocks:
            short = self._conv_bn(net, b.down, training, stats) if b.down is not None else net
            h = self._conv_bn(net, b.conv_a, training, stats)
            h = self._conv_bn(h, b.conv_b, training, stats)
            net = F.relu(h + short)
        pooled = net.mean(dim=2)
        if training and self.keep < 1.0:
            if mask is None:
                mask = torch.floor(self.keep + torch.rand(pooled.shape, generator=self.gen, dtype=pooled.dtype))
            pooled = pooled / self.keep * mask
        return pooled @ self.p[f"{self.spec.scope}/fc/weights"].reshape(self.spec.c_last, self.spec.num_classes)

    def train_step(self, wav: torch.Tensor, onehot: torch.Tensor, lr=0.1, momentum=0.9, weight_decay=1e-3, mask=None):
        """One session.run(train_op): front-end, forward, loss, backward, momentum, BN moving averages."""
        stats = {}
        with torch.no_grad():
            feat = self.mfcc(wav)
        logits = self.logits(feat, True, stats, mask)
        ce = -(onehot * F.log_softmax(logits, dim=1)).sum(dim=1).mean()
        l2 = sum((v * v).sum() / 2 for k, v in self.p.items() if "BatchNorm" not in k)
        total = ce + weight_decay * l2
        names = list(self.p)
        grads = torch.autograd.grad(total, [self.p[k] for k in names])
        with torch.no_grad():
            for k, g in zip(names, grads):
                self.slots[k].mul_(momentum).add_(g)
                self.p[k].sub_(lr * self.slots[k])
            for p, (mean, var) in stats.items():
                mm, mvv = self.mv[p + "/BatchNorm/moving_mean"], self.mv[p + "/BatchNorm/moving_variance"]
                mm.sub_((mm - mean) * (1.0 - O.BN_DECAY))
                mvv.sub_((mvv - var) * (1.0 - O.BN_DECAY))
        return float(total.detach()), float(ce.detach()), logits.detach()
