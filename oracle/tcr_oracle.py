"""CPU oracle for the TC-ResNet hot path (TEST INFRASTRUCTURE ONLY).

This is a NumPy restatement of the arithmetic that one ``session.run(train_op)``
of hyperconnect/TC-ResNet executes after the tf.data iterator has produced
``(wavs, labels)``: MFCC front-end, TCResNet8/14 forward, softmax-CE + L2 loss,
backward, SGD-momentum update, BN moving-average update.

PARITY UNPINNED.  The reference is pure Python on TensorFlow 1.13.1; TF is not
importable in this image and the reference ships no tests, golden vectors or
fixtures.  The arithmetic therefore follows the reference
call sites (cited per function below, paths relative to the reference root)
plus the published TF r1.13 op semantics.  Self-checks that stand in for the
missing pins live in tests/test_oracle.py: finite-difference gradient checks in
fp64, an independent PyTorch-autograd cross-check, scipy DCT identity, and an
independent NumPy port of the mel matrix.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product path (tc-resnet_b200/) never
does, and fails loudly when its CUDA library is missing.

Everything is parametric in ``dtype`` (np.float64 = ground truth, np.float32 =
tolerance calibration: what a faithful fp32 implementation can be expected to
reach).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# Front-end: datasets/preprocessors.py:64-96 (_log_mel_spectrogram), :183-194 (MFCC)
# --------------------------------------------------------------------------------------


def next_pow2(n: int) -> int:
    """fft_length default of tf.contrib.signal.stft: smallest power of two >= frame_length."""
    return 1 << (int(n) - 1).bit_length()


def num_frames(num_samples: int, window: int, stride: int) -> int:
    """tf.contrib.signal.frame(pad_end=False): T = 1 + (L - W) // S."""
    if num_samples < window:
        return 0
    return 1 + (num_samples - window) // stride


def hann_window_periodic(window: int, dtype=np.float64) -> np.ndarray:
    """tf.contrib.signal.hann_window(W, periodic=True) (window_fn default of stft,
    datasets/preprocessors.py:68).  TF: even = 1 - W % 2; n = W + periodic*even - 1;
    w = 0.5 - 0.5 cos(2 pi count / n), all in the signal dtype."""
    dt = np.dtype(dtype).type
    even = 1 - window % 2
    n = dt(window + even - 1)
    count = np.arange(window).astype(dtype)
    cos_arg = dt(2.0 * np.pi) * count / n
    return (dt(0.5) - dt(0.5) * np.cos(cos_arg)).astype(dtype)


def _hertz_to_mel(f, dtype):
    dt = np.dtype(dtype).type
    return dt(1127.0) * np.log(dt(1.0) + np.asarray(f, dtype=dtype) / dt(700.0))


def linear_to_mel_weight_matrix(num_mel_bins=64, num_spectrogram_bins=513, sample_rate=16000,
                                lower_edge_hertz=80.0, upper_edge_hertz=7600.0,
                                dtype=np.float64) -> np.ndarray:
    """tf.contrib.signal.linear_to_mel_weight_matrix as called at
    datasets/preprocessors.py:80-86.  HTK mel scale, DC bin zeroed, triangles
    max(0, min(lower_slope, upper_slope)); returns [bins, num_mel_bins]."""
    dt = np.dtype(dtype).type
    bands_to_zero = 1
    nyquist = dt(sample_rate) / dt(2.0)
    linear_frequencies = np.linspace(dt(0.0), nyquist, num_spectrogram_bins).astype(dtype)[bands_to_zero:]
    spectrogram_bins_mel = _hertz_to_mel(linear_frequencies, dtype)[:, None]
    edges = np.linspace(_hertz_to_mel(lower_edge_hertz, dtype), _hertz_to_mel(upper_edge_hertz, dtype),
                        num_mel_bins + 2).astype(dtype)
    lower_edge_mel = edges[None, :-2]
    center_mel = edges[None, 1:-1]
    upper_edge_mel = edges[None, 2:]
    lower_slopes = (spectrogram_bins_mel - lower_edge_mel) / (center_mel - lower_edge_mel)
    upper_slopes = (upper_edge_mel - spectrogram_bins_mel) / (upper_edge_mel - center_mel)
    w = np.maximum(dt(0.0), np.minimum(lower_slopes, upper_slopes))
    return np.pad(w, [[bands_to_zero, 0], [0, 0]]).astype(dtype)


def dct2_matrix(num_mel_bins=64, num_mfccs=40, dtype=np.float64) -> np.ndarray:
    """tf.contrib.signal.mfccs_from_log_mel_spectrograms (datasets/preprocessors.py:191-192):
    unnormalised DCT-II  y_k = 2 sum_n x_n cos(pi k (2n+1) / (2N)), scaled by rsqrt(2N),
    first num_mfccs coefficients.  Returned as a [N, num_mfccs] matrix D so y = x @ D."""
    n = np.arange(num_mel_bins, dtype=np.float64)[:, None]
    k = np.arange(num_mfccs, dtype=np.float64)[None, :]
    d = 2.0 * np.cos(np.pi * k * (2.0 * n + 1.0) / (2.0 * num_mel_bins)) / math.sqrt(2.0 * num_mel_bins)
    return d.astype(dtype)


def frame_signal(wav: np.ndarray, window: int, stride: int) -> np.ndarray:
    """[N, L] -> [N, T, W] strided frames, no end padding."""
    n, length = wav.shape
    t = num_frames(length, window, stride)
    idx = (np.arange(t) * stride)[:, None] + np.arange(window)[None, :]
    return wav[:, idx]


def power_spectrogram(wav: np.ndarray, window: int, stride: int, magnitude_squared=True,
                      dtype=np.float64) -> np.ndarray:
    """stft (periodic Hann, fft = next pow2, zero pad at end, no centring) then
    real(X conj X) (MFCC path) or |X| (log-mel path): datasets/preprocessors.py:67-77."""
    wav = np.asarray(wav, dtype=dtype)
    if wav.ndim == 3:
        wav = wav[..., 0]  # tf.squeeze(audio, -1)
    fft = next_pow2(window)
    frames = frame_signal(wav, window, stride) * hann_window_periodic(window, dtype)[None, None, :]
    spec = np.fft.rfft(frames, n=fft, axis=-1)
    if magnitude_squared:
        out = spec.real * spec.real + spec.imag * spec.imag
    else:
        out = np.abs(spec)
    return out.astype(dtype)


def log_mel_spectrogram(wav, window, stride, magnitude_squared=True, num_mel_bins=64,
                        sample_rate=16000, lower_edge_hertz=80.0, upper_edge_hertz=7600.0,
                        dtype=np.float64) -> np.ndarray:
    """datasets/preprocessors.py:64-96 -> [N, T, num_mel_bins]."""
    spec = power_spectrogram(wav, window, stride, magnitude_squared, dtype)
    mel_w = linear_to_mel_weight_matrix(num_mel_bins, spec.shape[-1], sample_rate,
                                        lower_edge_hertz, upper_edge_hertz, dtype)
    mel = spec @ mel_w
    return np.log(mel + np.dtype(dtype).type(1e-6)).astype(dtype)


def mfcc(wav, window, stride, num_mel_bins=64, num_mfccs=40, sample_rate=16000,
         lower_edge_hertz=80.0, upper_edge_hertz=7600.0, dtype=np.float64) -> np.ndarray:
    """MFCCPreprocessor._preprocess (datasets/preprocessors.py:183-194) without the
    trailing expand_dims: [N, L(,1)] -> [N, T, num_mfccs]."""
    lm = log_mel_spectrogram(wav, window, stride, True, num_mel_bins, sample_rate,
                             lower_edge_hertz, upper_edge_hertz, dtype)
    return (lm @ dct2_matrix(num_mel_bins, num_mfccs, dtype)).astype(dtype)


# --------------------------------------------------------------------------------------
# Network description: audio_nets/tc_resnet.py:6-70
# --------------------------------------------------------------------------------------


def same_padding(length: int, k: int, s: int) -> Tuple[int, int, int]:
    """TF 'SAME': out = ceil(L/s); total = max((out-1)s + k - L, 0); left = total // 2."""
    out = -(-length // s)
    total = max((out - 1) * s + k - length, 0)
    left = total // 2
    return out, left, total - left


@dataclass
class ConvSpec:
    name: str          # TF scope under <Net>/, e.g. "block0/conv0_0"
    cin: int
    cout: int
    k: int
    stride: int
    t_in: int
    t_out: int
    pad_left: int
    pad_right: int
    relu: bool         # activation after BN (False only for conv{i}_1)


@dataclass
class BlockSpec:
    index: int
    down: Optional[ConvSpec]
    conv_a: ConvSpec   # conv{i}_0
    conv_b: ConvSpec   # conv{i}_1


@dataclass
class NetSpec:
    scope: str                     # "TCResNet8" / "TCResNet14"
    t_in: int
    f_in: int
    num_classes: int
    conv0: ConvSpec
    blocks: List[BlockSpec]
    c_last: int
    t_last: int
    # trainable variable table in tf.trainable_variables() order
    var_names: List[str] = field(default_factory=list)
    var_shapes: Dict[str, Tuple[int, ...]] = field(default_factory=dict)
    bn_layers: List[str] = field(default_factory=list)   # conv scopes owning a BatchNorm

    def convs(self) -> List[ConvSpec]:
        out = [self.conv0]
        for b in self.blocks:
            if b.down is not None:
                out.append(b.down)
            out += [b.conv_a, b.conv_b]
        return out


def build_spec(model: str = "TCResNet8", width_multiplier: float = 1.0, t_in: int = 49, f_in: int = 40,
               num_classes: int = 12) -> NetSpec:
    """audio_nets/tc_resnet.py:57-70 channel plans, :6-54 topology."""
    if model in ("TCResNet8", "TCResNet8Model", 8):
        scope, plan = "TCResNet8", [16, 24, 32, 48]
    elif model in ("TCResNet14", "TCResNet14Model", 14):
        scope, plan = "TCResNet14", [16, 24, 24, 32, 32, 48, 48]
    else:
        raise ValueError(f"unknown model {model}")
    plan = [int(x * width_multiplier) for x in plan]

    def mk(name, cin, cout, k, s, t, relu=True):
        t_out, pl, pr = same_padding(t, k, s)
        return ConvSpec(name, cin, cout, k, s, t, t_out, pl, pr, relu)

    conv0 = mk("conv0", f_in, plan[0], 3, 1, t_in)
    blocks = []
    c, t = plan[0], conv0.t_out
    for i, n in enumerate(plan[1:]):
        if n != c:
            stride = 2
            down = mk(f"block{i}/down", c, n, 1, 2, t)
        else:
            stride = 1
            down = None
        ca = mk(f"block{i}/conv{i}_0", c, n, 9, stride, t)
        cb = mk(f"block{i}/conv{i}_1", n, n, 9, 1, ca.t_out, relu=False)
        blocks.append(BlockSpec(i, down, ca, cb))
        c, t = n, cb.t_out
    spec = NetSpec(scope, t_in, f_in, num_classes, conv0, blocks, c, t)
    for cv in spec.convs():
        spec.var_names.append(f"{scope}/{cv.name}/weights")
        spec.var_shapes[spec.var_names[-1]] = (cv.k, 1, cv.cin, cv.cout)
        for p in ("beta", "gamma"):
            spec.var_names.append(f"{scope}/{cv.name}/BatchNorm/{p}")
            spec.var_shapes[spec.var_names[-1]] = (cv.cout,)
        spec.bn_layers.append(cv.name)
    for fc, n_out in (("fc", num_classes), ("fc2", 2)):
        spec.var_names.append(f"{scope}/{fc}/weights")
        spec.var_shapes[spec.var_names[-1]] = (1, 1, c, n_out)
    return spec


def count_trainable(spec: NetSpec) -> int:
    return int(sum(int(np.prod(spec.var_shapes[n])) for n in spec.var_names))


def forward_flops(spec: NetSpec) -> int:
    """2 FLOP per MAC, convs + fc + fc2 (SURVEY.md layer table)."""
    f = 0
    for cv in spec.convs():
        f += 2 * cv.t_out * cv.k * cv.cin * cv.cout
    f += 2 * spec.c_last * spec.num_classes + 2 * spec.c_last * 2
    return f


def init_variables(spec: NetSpec, seed: int = 0, dtype=np.float64):
    """TCResNet_arg_scope (audio_nets/tc_resnet.py:102-123): Xavier-uniform weights
    U(+-sqrt(6/(fan_in+fan_out))), gamma=1, beta=0, moving mean/var = 0/1.
    TF's RNG is not reproducible, so parity always *injects* these values."""
    rng = np.random.RandomState(seed)
    params, moving = {}, {}
    for name in spec.var_names:
        shape = spec.var_shapes[name]
        if name.endswith("/weights"):
            k, _, cin, cout = shape
            limit = math.sqrt(6.0 / (k * cin + k * cout))
            params[name] = rng.uniform(-limit, limit, size=shape).astype(dtype)
        elif name.endswith("/gamma"):
            params[name] = np.ones(shape, dtype)
        else:
            params[name] = np.zeros(shape, dtype)
    for layer in spec.bn_layers:
        c = spec.var_shapes[f"{spec.scope}/{layer}/weights"][3]
        moving[f"{spec.scope}/{layer}/BatchNorm/moving_mean"] = np.zeros((c,), dtype)
        moving[f"{spec.scope}/{layer}/BatchNorm/moving_variance"] = np.ones((c,), dtype)
    return params, moving


# --------------------------------------------------------------------------------------
# Layer primitives (slim.conv2d NHWC [k,1] cross-correlation, fused batch norm)
# --------------------------------------------------------------------------------------


def _im2col(x: np.ndarray, cv: ConvSpec) -> np.ndarray:
    n = x.shape[0]
    xp = np.pad(x, [[0, 0], [cv.pad_left, cv.pad_right], [0, 0]])
    idx = (np.arange(cv.t_out) * cv.stride)[:, None] + np.arange(cv.k)[None, :]
    cols = xp[:, idx, :]                      # [N, T', K, Cin]
    return cols.reshape(n * cv.t_out, cv.k * cv.cin)


def conv_forward(x, w, cv: ConvSpec):
    """y[n,t,co] = sum_{k,ci} xpad[n, s t + k, ci] w[k,0,ci,co]; no bias (biases_initializer=None)."""
    cols = _im2col(x, cv)
    y = cols @ w.reshape(cv.k * cv.cin, cv.cout)
    return y.reshape(x.shape[0], cv.t_out, cv.cout), cols


def conv_backward(dy, cols, w, cv: ConvSpec, need_dx=True):
    n = dy.shape[0]
    dy2 = dy.reshape(n * cv.t_out, cv.cout)
    dw = (cols.T @ dy2).reshape(cv.k, 1, cv.cin, cv.cout)
    dx = None
    if need_dx:
        dcols = (dy2 @ w.reshape(cv.k * cv.cin, cv.cout).T).reshape(n, cv.t_out, cv.k, cv.cin)
        dxp = np.zeros((n, cv.t_in + cv.pad_left + cv.pad_right, cv.cin), dy.dtype)
        for k in range(cv.k):
            dxp[:, k:k + cv.stride * cv.t_out:cv.stride, :] += dcols[:, :, k, :]
        dx = dxp[:, cv.pad_left:cv.pad_left + cv.t_in, :]
    return dx, dw


BN_EPS = 1e-3      # slim.batch_norm default epsilon
BN_DECAY = 0.997   # audio_nets/tc_resnet.py:107


def bn_forward_train(y, gamma, beta, eps=BN_EPS):
    """tf.nn.fused_batch_norm(is_training=True): per-channel mean and BIASED variance over
    N*T' for normalisation; the variance handed to the moving average is UNBIASED."""
    dt = y.dtype.type
    m = y.shape[0] * y.shape[1]
    mean = y.mean(axis=(0, 1))
    xc = y - mean
    var = (xc * xc).mean(axis=(0, 1))
    rstd = dt(1.0) / np.sqrt(var + dt(eps))
    xhat = xc * rstd
    z = xhat * gamma + beta
    var_unbiased = var * dt(m / max(m - 1, 1))
    return z, (xhat, rstd, mean, var, var_unbiased)


def bn_forward_eval(y, gamma, beta, moving_mean, moving_var, eps=BN_EPS):
    dt = y.dtype.type
    rstd = dt(1.0) / np.sqrt(moving_var + dt(eps))
    return (y - moving_mean) * rstd * gamma + beta


def bn_backward_train(dz, xhat, rstd, gamma):
    """FusedBatchNormGrad (training)."""
    m = dz.shape[0] * dz.shape[1]
    dbeta = dz.sum(axis=(0, 1))
    dgamma = (dz * xhat).sum(axis=(0, 1))
    dy = (gamma * rstd) * (dz - dbeta / m - xhat * (dgamma / m))
    return dy, dgamma, dbeta


# --------------------------------------------------------------------------------------
# Forward / loss / backward / optimizer
# --------------------------------------------------------------------------------------


def _conv_bn(x, cv: ConvSpec, spec: NetSpec, params, moving, is_training, cache, forced_masks=None):
    p = f"{spec.scope}/{cv.name}"
    y, cols = conv_forward(x, params[p + "/weights"], cv)
    gamma, beta = params[p + "/BatchNorm/gamma"], params[p + "/BatchNorm/beta"]
    if is_training:
        z, bn = bn_forward_train(y, gamma, beta)
    else:
        z = bn_forward_eval(y, gamma, beta, moving[p + "/BatchNorm/moving_mean"],
                            moving[p + "/BatchNorm/moving_variance"])
        bn = None
    # forced_masks: the ReLU decisions of ANOTHER run of the same step (tests pass the CUDA path's), so that activations within
    # round-off of zero do not flip between the two and the comparison sees arithmetic differences only
    mask = forced_masks[cv.name] if (forced_masks is not None and cv.relu) else (z > 0)
    a = np.where(mask, z, 0).astype(z.dtype) if cv.relu else z
    cache[cv.name] = dict(x=x, cols=cols, y=y, z=z, bn=bn, a=a, mask=mask)
    return a


def forward(spec: NetSpec, params, moving, feat, is_training: bool,
            keep_prob: float = 1.0, dropout_mask: Optional[np.ndarray] = None, forced_masks=None):
    """tc_resnet (audio_nets/tc_resnet.py:6-54) on features [N, T, F] (== [N,T,1,F] NHWC).
    dropout_mask: optional {0,1} array [N, C_last] (floor(keep + U)); TF's RNG is not
    reproducible so parity uses keep_prob=1.0 or an injected mask."""
    dt = feat.dtype.type
    cache: Dict[str, dict] = {}
    net = _conv_bn(feat, spec.conv0, spec, params, moving, is_training, cache, forced_masks)
    for b in spec.blocks:
        if b.down is not None:
            short = _conv_bn(net, b.down, spec, params, moving, is_training, cache, forced_masks)  # BN + ReLU (quirk)
        else:
            short = net
        h = _conv_bn(net, b.conv_a, spec, params, moving, is_training, cache, forced_masks)
        h = _conv_bn(h, b.conv_b, spec, params, moving, is_training, cache, forced_masks)
        pre = h + short
        bmask = forced_masks[f"block{b.index}"] if forced_masks is not None else (pre > 0)
        net = np.where(bmask, pre, 0).astype(pre.dtype)
        cache[f"block{b.index}"] = dict(short=short, out=net, mask=bmask)
    pooled = net.mean(axis=1)                                  # avg_pool over full [T',1]
    if is_training and keep_prob < 1.0:
        if dropout_mask is None:
            raise ValueError("training with keep_prob < 1 needs an injected dropout mask")
        dropped = pooled / dt(keep_prob) * dropout_mask.astype(feat.dtype)
    else:
        dropped = pooled
    wfc = params[f"{spec.scope}/fc/weights"].reshape(spec.c_last, spec.num_classes)
    logits = dropped @ wfc
    cache["head"] = dict(pooled=pooled, dropped=dropped, mask=dropout_mask, keep=keep_prob)
    return logits, cache


def softmax(logits):
    z = logits - logits.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)


def losses(spec: NetSpec, params, logits, onehot, weight_decay: float, label_smoothing: float = 0.0):
    """AudioNetModel.build_loss (factory/audio_nets.py:161-183)."""
    dt = logits.dtype.type
    onehot = onehot.astype(logits.dtype)
    if label_smoothing > 0:
        onehot = onehot * dt(1.0 - label_smoothing) + dt(label_smoothing / spec.num_classes)
    z = logits - logits.max(axis=1, keepdims=True)
    logp = z - np.log(np.exp(z).sum(axis=1, keepdims=True))
    model_loss = -(onehot * logp).sum(axis=1).mean()
    l2 = dt(0.0)
    for name in spec.var_names:
        if "BatchNorm" not in name:
            l2 = l2 + (params[name] ** 2).sum() / dt(2.0)
    total = model_loss + dt(weight_decay) * l2
    return total, model_loss


def backward(spec: NetSpec, params, cache, logits, onehot, weight_decay: float, label_smoothing: float = 0.0):
    """Gradient of total_loss w.r.t. every trainable (slim.learning.create_train_op,
    helper/trainer.py:199-211): g = dCE/dv + weight_decay * v for non-BN variables."""
    dt = logits.dtype.type
    n = logits.shape[0]
    onehot = onehot.astype(logits.dtype)
    if label_smoothing > 0:
        onehot = onehot * dt(1.0 - label_smoothing) + dt(label_smoothing / spec.num_classes)
    grads: Dict[str, np.ndarray] = {}
    sc = spec.scope
    probs = softmax(logits)
    dlogits = (probs * onehot.sum(axis=1, keepdims=True) - onehot) / dt(n)
    head = cache["head"]
    wfc = params[f"{sc}/fc/weights"].reshape(spec.c_last, spec.num_classes)
    grads[f"{sc}/fc/weights"] = (head["dropped"].T @ dlogits).reshape(1, 1, spec.c_last, spec.num_classes)
    ddropped = dlogits @ wfc.T
    if head["mask"] is not None and head["keep"] < 1.0:
        dpooled = ddropped / dt(head["keep"]) * head["mask"].astype(logits.dtype)
    else:
        dpooled = ddropped
    dnet = np.repeat(dpooled[:, None, :], spec.t_last, axis=1) / dt(spec.t_last)

    def conv_bn_back(cv: ConvSpec, da, need_dx=True):
        c = cache[cv.name]
        p = f"{sc}/{cv.name}"
        dz = da * c["mask"] if cv.relu else da
        xhat, rstd = c["bn"][0], c["bn"][1]
        dy, dgamma, dbeta = bn_backward_train(dz, xhat, rstd, params[p + "/BatchNorm/gamma"])
        dx, dw = conv_backward(dy, c["cols"], params[p + "/weights"], cv, need_dx)
        grads[p + "/weights"] = dw
        grads[p + "/BatchNorm/gamma"] = dgamma
        grads[p + "/BatchNorm/beta"] = dbeta
        c["dz"], c["dy"] = dz, dy
        return dx

    for b in reversed(spec.blocks):
        g = dnet * cache[f"block{b.index}"]["mask"]
        cache[f"block{b.index}"]["g"] = g
        dh = conv_bn_back(b.conv_b, g)
        dx = conv_bn_back(b.conv_a, dh)
        if b.down is not None:
            dx = dx + conv_bn_back(b.down, g)
        else:
            dx = dx + g
        dnet = dx
    conv_bn_back(spec.conv0, dnet, need_dx=False)
    grads[f"{sc}/fc2/weights"] = np.zeros_like(params[f"{sc}/fc2/weights"])   # dead head (quirk)
    for name in spec.var_names:
        if "BatchNorm" not in name:
            grads[name] = grads[name] + dt(weight_decay) * params[name]
    return grads


def piecewise_constant(step: int, boundaries, values) -> float:
    """tf.train.piecewise_constant (helper/trainer.py:135): values[i] while step <= boundaries[i]."""
    for b, v in zip(boundaries, values):
        if step <= b:
            return v
    return values[-1]


def train_step(spec: NetSpec, params, moving, slots, feat, onehot, lr: float, momentum: float = 0.9,
               weight_decay: float = 1e-3, keep_prob: float = 1.0, dropout_mask=None,
               label_smoothing: float = 0.0, bn_decay: float = BN_DECAY, forced_masks=None):
    """One session.run(train_op): forward (batch-stat BN), loss, backward, BN moving-average
    update (UPDATE_OPS), MomentumOptimizer: m <- mom*m + g ; v <- v - lr*m
    (helper/trainer.py:171-222).  Returns new (params, moving, slots), losses and extras."""
    dt = feat.dtype.type
    logits, cache = forward(spec, params, moving, feat, True, keep_prob, dropout_mask, forced_masks)
    total, model_loss = losses(spec, params, logits, onehot, weight_decay, label_smoothing)
    grads = backward(spec, params, cache, logits, onehot, weight_decay, label_smoothing)
    new_params, new_slots, new_moving = {}, {}, dict(moving)
    for name in spec.var_names:
        m = slots[name] * dt(momentum) + grads[name]
        new_slots[name] = m
        new_params[name] = params[name] - dt(lr) * m
    one_minus = dt(1.0 - bn_decay)
    for cv in spec.convs():
        p = f"{spec.scope}/{cv.name}/BatchNorm/"
        _, _, mean, _, var_unb = cache[cv.name]["bn"]
        new_moving[p + "moving_mean"] = moving[p + "moving_mean"] - (moving[p + "moving_mean"] - mean) * one_minus
        new_moving[p + "moving_variance"] = (moving[p + "moving_variance"]
                                             - (moving[p + "moving_variance"] - var_unb) * one_minus)
    return new_params, new_moving, new_slots, dict(total_loss=total, model_loss=model_loss,
                                                   logits=logits, grads=grads, cache=cache)


def zeros_like_vars(spec: NetSpec, dtype=np.float64):
    return {n: np.zeros(spec.var_shapes[n], dtype) for n in spec.var_names}


# --------------------------------------------------------------------------------------
# Flat-buffer helpers (the C ABI passes flat fp32 buffers; order = spec.var_names)
# --------------------------------------------------------------------------------------


def flatten_vars(spec: NetSpec, d: Dict[str, np.ndarray], dtype=np.float32) -> np.ndarray:
    return np.concatenate([np.asarray(d[n], dtype).ravel() for n in spec.var_names])


def unflatten_vars(spec: NetSpec, flat: np.ndarray, dtype=np.float64) -> Dict[str, np.ndarray]:
    out, off = {}, 0
    for n in spec.var_names:
        sz = int(np.prod(spec.var_shapes[n]))
        out[n] = np.asarray(flat[off:off + sz], dtype).reshape(spec.var_shapes[n])
        off += sz
    return out


def moving_names(spec: NetSpec) -> List[str]:
    out = []
    for layer in spec.bn_layers:
        out.append(f"{spec.scope}/{layer}/BatchNorm/moving_mean")
        out.append(f"{spec.scope}/{layer}/BatchNorm/moving_variance")
    return out


def flatten_moving(spec: NetSpec, d, dtype=np.float32) -> np.ndarray:
    return np.concatenate([np.asarray(d[n], dtype).ravel() for n in moving_names(spec)])


def unflatten_moving(spec: NetSpec, flat, dtype=np.float64):
    out, off = {}, 0
    for n in moving_names(spec):
        layer = n.split("/BatchNorm/")[0]
        c = spec.var_shapes[layer + "/weights"][3]
        out[n] = np.asarray(flat[off:off + c], dtype)
        off += c
    return out


def cast_vars(d, dtype):
    return {k: np.asarray(v, dtype) for k, v in d.items()}


def synthetic_batch(n: int, num_samples: int = 16000, num_classes: int = 12, seed_wav: int = 1234,
                    seed_label: int = 4321, adversarial: bool = False):
    """wav = U(-1,1) fp32 [N, L]; labels randint -> one-hot fp32.
    adversarial=True replaces clip 0 by silence and clip 1 by a full-scale square wave."""
    rng = np.random.RandomState(seed_wav)
    wav = rng.uniform(-1.0, 1.0, size=(n, num_samples)).astype(np.float32)
    if adversarial and n >= 2:
        wav[0] = 0.0
        wav[1] = np.where((np.arange(num_samples) // 40) % 2 == 0, 1.0, -1.0).astype(np.float32)
    lab = np.random.RandomState(seed_label).randint(0, num_classes, size=(n,))
    onehot = np.zeros((n, num_classes), np.float32)
    onehot[np.arange(n), lab] = 1.0
    return wav, onehot
