"""Parity cases shared by the emulator (CPU) and GPU tests: CUDA path through the C ABI vs the fp64 oracle.

Tolerance (BASELINE.json north_star): max_i |x_cuda - x_ref| / max_i |x_ref| <= 1e-4 for logits
against the fp64 oracle, arg-max class identical on every row whose top-2 margin exceeds that error; the same
rule for gradients / post-step state, floored by the fp32 oracle's own distance to fp64.
"""
from __future__ import annotations

import numpy as np

from oracle import tcr_oracle as O
from tcr_harness import Engine, rel_err

TOL_FEATURES = 2e-5
TOL_LOGITS = 1e-4
TOL_STATE = 1e-4


def perturbed_variables(spec, seed=0):
    """Xavier weights (NumPy seed 0) with non-trivial gamma/beta/moving stats so BN paths are exercised."""
    params, moving = O.init_variables(spec, seed)
    rng = np.random.RandomState(seed + 1)
    for k in params:
        if k.endswith("gamma"):
            params[k] = 1.0 + 0.2 * rng.randn(*params[k].shape)
        if k.endswith("beta"):
            params[k] = 0.1 * rng.randn(*params[k].shape)
    for k in moving:
        moving[k] = (rng.rand(*moving[k].shape) + 0.5) if k.endswith("variance") else 0.1 * rng.randn(*moving[k].shape)
    return params, moving


def cuda_relu_masks(eng, spec, n):
    """The ReLU decisions the CUDA path took in its last training step, rebuilt from its workspace (pre-BN outputs `y:<conv>`,
    the BN tables `bnf:<conv>` and the block outputs `out:<block>`): bn(y) = fma(y - mean, scale, beta) > 0.  Forcing them on the
    oracle removes mask flips of activations within round-off of zero from the gradient comparison (they dominate it at
    batch 512 / 1024: a flipped unit changes its whole fan-in's gradient), so the bound on gradients can be the plain 1e-4."""
    masks = {}
    for cv in spec.convs():
        if not cv.relu:
            continue
        y = eng.workspace(f"y:{cv.name}").reshape(-1, cv.t_out, cv.cout)[:n]
        t = eng.workspace(f"bnf:{cv.name}").reshape(4, cv.cout)
        d = (y - t[0]).astype(np.float32).astype(np.float64)            # the subtraction rounds to fp32, the fma does not
        masks[cv.name] = (d * t[2].astype(np.float64) + t[3].astype(np.float64)) > 0
    for b in spec.blocks:
        masks[f"block{b.index}"] = eng.workspace(f"out:block{b.index}").reshape(-1, b.conv_b.t_out, b.conv_b.cout)[:n] > 0
    return masks


def check_argmax(logits, ref_logits, err_abs):
    ref_sorted = np.sort(ref_logits, axis=1)
    margin = ref_sorted[:, -1] - ref_sorted[:, -2]
    decided = margin > 2 * err_abs
    assert np.array_equal(np.argmax(logits, 1)[decided], np.argmax(ref_logits, 1)[decided])
    return int((~decided).sum())


def run_case(backend, model="TCResNet8", wm=1.0, window=640, stride=320, n=4, keep=1.0, ls=0.0, use_wav=True,
             steps=1, adversarial=True, max_batch=None, check_f32_floor=False, force_masks=False):
    t = O.num_frames(16000, window, stride)
    spec = O.build_spec(model, wm, t)
    eng = Engine(backend, model=int(model[len("TCResNet"):]), width_multiplier=wm, max_batch=max_batch or max(n, 8),
                 dropout_keep_prob=keep, window_size_samples=window, window_stride_samples=stride, label_smoothing=ls)
    try:
        assert eng.info.frames == t and eng.info.num_trainable == O.count_trainable(spec)
        assert eng.info.forward_flops_per_utt == O.forward_flops(spec)
        tab = eng.param_table()
        assert [d["name"] for d in tab if d["kind"] < 3] == spec.var_names
        assert [d["name"] for d in tab if d["kind"] >= 3] == O.moving_names(spec)
        params, moving = perturbed_variables(spec)
        rng = np.random.RandomState(7)
        slots = {k: 0.01 * rng.randn(*v.shape) for k, v in params.items()}
        wav, onehot = O.synthetic_batch(n, adversarial=adversarial)
        feat = O.mfcc(wav, window, stride)
        mask = (rng.rand(n, spec.c_last) < keep).astype(np.float32) if keep < 1 else None
        lr, mom, wd = 0.1, 0.9, 1e-3
        report = {}

        # front-end
        f_cuda = eng.mfcc(wav)
        report["features"] = rel_err(f_cuda, feat)
        assert report["features"] <= TOL_FEATURES, report

        inp = wav if use_wav else feat.astype(np.float32)
        pf, mf, sf = O.flatten_vars(spec, params), O.flatten_moving(spec, moving), O.flatten_vars(spec, slots)

        # evaluate_audio.py path: moving statistics, dropout identity
        ev = eng.forward(inp, pf, mf, is_features=not use_wav, onehot_np=onehot, weight_decay=wd)
        ref_logits, _ = O.forward(spec, params, moving, feat, False)
        ref_total, ref_model = O.losses(spec, params, ref_logits, onehot, wd, ls)
        report["eval_logits"] = rel_err(ev["logits"], ref_logits)
        assert report["eval_logits"] <= TOL_LOGITS, report
        check_argmax(ev["logits"], ref_logits, report["eval_logits"] * np.abs(ref_logits).max())
        np.testing.assert_allclose(ev["probs"], O.softmax(ref_logits), rtol=0, atol=2e-5)
        assert abs(ev["losses"][0] - ref_total) <= 1e-4 * abs(ref_total) + 1e-6
        assert abs(ev["losses"][1] - ref_model) <= 1e-4 * abs(ref_model) + 1e-6

        # training steps.  Each step is compared against the oracle started from the CUDA path's OWN previous state
        # (re-synchronised), so the check is per-step parity + correct carrying of state between calls, not the
        # divergence of two chaotic lr=0.1 trajectories (ReLU masks flip under 1e-6 perturbations).
        pf_c, mf_c, sf_c = pf, mf, sf
        for step in range(steps):
            p = O.unflatten_vars(spec, pf_c)
            mv = O.unflatten_moving(spec, mf_c)
            sl = O.unflatten_vars(spec, sf_c)
            ts = eng.train_step(inp, onehot, pf_c, sf_c, mf_c, lr, mom, wd, is_features=not use_wav, mask_np=mask, seed=step)
            forced = cuda_relu_masks(eng, spec, n) if force_masks else None
            p1, mv1, sl1, ref = O.train_step(spec, p, mv, sl, feat, onehot, lr, mom, wd, keep, mask, ls, forced_masks=forced)
            pf_c, mf_c, sf_c = ts["params"], ts["moving"], ts["slots"]
            report[f"train_logits{step}"] = rel_err(ts["logits"], ref["logits"])
            report[f"grads{step}"] = rel_err(ts["grads"], O.flatten_vars(spec, ref["grads"], np.float64))
            report[f"params{step}"] = rel_err(ts["params"], O.flatten_vars(spec, p1, np.float64))
            report[f"slots{step}"] = rel_err(ts["slots"], O.flatten_vars(spec, sl1, np.float64))
            report[f"moving{step}"] = rel_err(ts["moving"], O.flatten_moving(spec, mv1, np.float64))
            floor = 0.0
            if check_f32_floor:                      # every step: the fp32 NumPy oracle's own distance to fp64 on this input
                p32, mv32, sl32 = (O.cast_vars(d, np.float32) for d in (p, mv, sl))
                _, _, _, r32 = O.train_step(spec, p32, mv32, sl32, feat.astype(np.float32), onehot, lr, mom, wd, keep, mask, ls)
                floor = 4 * rel_err(O.flatten_vars(spec, r32["grads"], np.float64), O.flatten_vars(spec, ref["grads"], np.float64))
                report[f"grads_f32_oracle_floor{step}"] = floor / 4
            assert report[f"train_logits{step}"] <= TOL_LOGITS, report
            assert report[f"grads{step}"] <= max(TOL_STATE, floor), report
            assert report[f"params{step}"] <= max(TOL_STATE, floor), report
            assert report[f"slots{step}"] <= max(TOL_STATE, floor), report
            assert report[f"moving{step}"] <= TOL_STATE, report
            assert abs(ts["losses"][0] - ref["total_loss"]) <= 1e-4 * abs(ref["total_loss"]) + 1e-6
            assert abs(ts["losses"][1] - ref["model_loss"]) <= 1e-4 * abs(ref["model_loss"]) + 1e-6
        return report
    finally:
        eng.close()


def spectral_lines():
    """Seven full-scale sines (on and between bin centres; 2296.875 Hz = bin 147 has the smallest falling mel weight) over a
    -60 dB noise floor: the worst case of the frame-pair front-end kernel's run-based mel stage."""
    t = np.arange(16000) / 16000.0
    rng = np.random.default_rng(3)
    return np.stack([np.sin(2 * np.pi * f * t) for f in (100.0, 437.3, 2296.875, 2345.6, 3999.0, 6999.9, 7590.0)]) \
        + 1e-3 * rng.uniform(-1, 1, (7, 16000))
