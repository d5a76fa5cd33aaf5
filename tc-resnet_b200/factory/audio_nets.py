"""Model builders with the reference's names and attribute surface (factory/audio_nets.py:19-183, 362-409),
executing on the CUDA engine instead of building a TF graph.

`train_audio.py` / `evaluate_audio.py` look the class up by NAME (`eval(f"audio_nets.{args.model}")`,
train_audio.py:32), call `Model(args, dataset)`, then `model.build(wavs, labels, is_training)` and hand the model
to the trainer / evaluator, which touch only the attributes set here and `session.run`.  TCResNet8Model and
TCResNet14Model are on the accelerated path; the other entries of `_available_nets` (baselines the north star
does not name) are registered so the CLI accepts them but raise NotImplementedError when built.
"""
from __future__ import annotations

from typing import Dict, Optional, Set

import numpy as np
import torch

from ..common import utils
from ..datasets import preprocessor_factory
from ..engine import Engine
from ..runtime import InvalidArgumentError, Node
from .base import TFModel

_available_nets = [
    "KWSModel", "Res8Model", "Res8NarrowModel", "Res15Model", "Res15NarrowModel", "DSCNNSModel", "DSCNNMModel",
    "DSCNNLModel", "TCResNet8Model", "TCResNet14Model", "ResNet2D8Model", "ResNet2D8PoolModel",
]


class AudioNetModel(TFModel):
    engine_model: Optional[str] = None      # set by the accelerated subclasses

    def __init__(self, args, dataset=None):
        self.log = utils.get_logger("AudioNetModel")
        self.args, self.dataset = args, dataset
        self.engine: Optional[Engine] = None
        self.global_step = 0
        self.optimizer = dict(kind="mom", momentum=0.9)            # filled in by the trainer (build_optimizer)
        self.lr_schedule = lambda step: getattr(args, "learning_rate", 1e-4)
        self.var_names_to_values = None                            # in-memory weight injection hook (trainer.py:145-154)
        self.endpoints_loss: Dict[str, Node] = {}
        self._last: Dict[str, np.ndarray] = {}
        self._counts = None                                        # device int64 [classes^2 + 2] (tcr_eval_accumulate)

    # ------------------------------------------------------------------ graph-construction API of the reference
    def build(self, wavs, labels, is_training):
        self._audio_original, self.labels, self.is_training = wavs, labels, is_training
        self.preprocess_input()
        self.inputs, self.logits, self._outputs, self.endpoints = self.build_output(
            self.audio, self.is_training, self.args.output_name)
        self._total_loss, self._model_loss, self.endpoints_loss = self.build_loss(self.logits, self.outputs, self.labels)
        self.total_params = self.engine.num_trainable + self.engine.num_moving
        self.log.info(f"{type(self).__name__}: {self.engine.num_trainable} trainable parameters, "
                      f"{self.engine.info.forward_flops_per_utt / 1e6:.3f} MFLOP forward per utterance")
        if self.dataset is not None and getattr(self.dataset, "session", None) is not None:
            self.dataset.session.bind(self)

    def preprocess_input(self, for_deploy=False):
        a = self.args
        window = int(a.sample_rate * a.window_size_ms / 1000)
        stride = int(a.sample_rate * a.window_stride_ms / 1000)
        pre = preprocessor_factory.factory(preprocess_method=a.preprocess_method, scope="input/audio/preprocessing",
                                           preprocessed_node_name="input/audio/preprocessed")
        self._audio = pre.preprocess(self._audio_original, window_size_samples=window, window_stride_samples=stride,
                                     for_deploy=for_deploy, **vars(a))
        self._audio.name = "audio"
        self.log.info(f"Update height/width to {self._audio.shape}")
        a.height, a.width, a.channels = self._audio.shape[1:4]
        self.input_preprocessors_for_tflite = [pre]

    def build_output(self, inputs, is_training, output_name):
        logits, endpoints = self.build_inference(inputs, is_training=is_training)
        return inputs, logits, Node("outputs", [None, self.args.num_classes]), endpoints

    def build_inference(self, inputs, is_training=True):
        if self.engine_model is None:
            raise NotImplementedError(f"{type(self).__name__} is a baseline outside the accelerated TC-ResNet path")
        a = self.args
        self.engine = Engine(model=self.engine_model, width_multiplier=a.width_multiplier, num_classes=a.num_classes,
                             sample_rate=a.sample_rate, clip_duration_ms=a.clip_duration_ms,
                             window_size_ms=a.window_size_ms, window_stride_ms=a.window_stride_ms,
                             num_mel_bins=a.num_mel_bins, num_mfccs=a.num_mfccs, lower_edge_hertz=a.lower_edge_hertz,
                             upper_edge_hertz=a.upper_edge_hertz, preprocess_method=a.preprocess_method,
                             max_batch=a.batch_size, dropout_keep_prob=a.dropout_keep_prob,
                             label_smoothing=getattr(a, "label_smoothing", 0.0))
        self.params, self.slots, self.moving = self.engine.new_variables(seed=getattr(a, "seed", 0) or 0)
        if getattr(a, "data_parallel", False):
            self.engine.attach_process_group()
        n = a.batch_size
        dev = self.engine.device
        self._h_wav = torch.empty(n, self.engine.cfg.clip_samples, dtype=torch.float32).pin_memory()
        self._h_hot = torch.empty(n, a.num_classes, dtype=torch.float32).pin_memory()
        self._d_wav, self._d_hot = self._h_wav.to(dev), self._h_hot.to(dev)
        return Node("logits", [None, a.num_classes]), {"ranges": Node("ranges")}

    def build_loss(self, logits, scores, labels):
        return Node("total_loss", []), Node("model_loss", []), {}

    def build_deployable_model(self, include_preprocess=True):
        raise NotImplementedError("TFLite freezing (freeze.py) is outside the accelerated path")

    # ------------------------------------------------------------------ attribute surface
    @property
    def model_loss(self):
        return self._model_loss

    @property
    def total_loss(self):
        return self._total_loss

    @property
    def audio_original(self):
        return self._audio_original

    @property
    def audio(self):
        return self._audio

    @property
    def outputs(self):
        return self._outputs

    # ------------------------------------------------------------------ variables by TF name
    def get_variables(self, with_slots=True) -> Dict[str, np.ndarray]:
        out = self.engine.variables_to_dict(self.params, self.moving)
        if with_slots:
            for k, v in self.engine.variables_to_dict(self.slots).items():
                out[k + "/Momentum"] = v
        return out

    def set_variables(self, values: Dict[str, np.ndarray], strict=True):
        self.engine.variables_from_dict(values, self.params, self.moving, strict=strict)
        slots = {k[:-len("/Momentum")]: v for k, v in values.items() if k.endswith("/Momentum")}
        if slots:
            self.engine.variables_from_dict(slots, self.slots, None, strict=False)

    def _dropout_seed(self):
        """Counter-RNG seed of this step: distinct per rank under data parallelism (the kernel indexes utterances locally)."""
        world = getattr(self.engine, "world_size", 1)
        rank = getattr(self.dataset, "rank", 0) if world > 1 else 0
        return self.global_step * world + rank

    def read_metric_counts(self, reset: bool = True) -> np.ndarray:
        """Confusion matrix / top-5 hits / sample count accumulated by the forward passes since the last reset (one D2H read)."""
        if self._counts is None:
            return np.zeros(self.args.num_classes ** 2 + 2, np.int64)
        out = self._counts.cpu().numpy()
        if reset:
            self._counts.zero_()
        return out

    # ------------------------------------------------------------------ one session.run
    def execute(self, names: Set[str], feed) -> Dict[str, object]:
        wav_np, hot_np = self.dataset.next_batch()
        n = wav_np.shape[0]
        if n != self.args.batch_size or not np.isfinite(wav_np).all():
            raise InvalidArgumentError(f"bad batch: {wav_np.shape}")
        self._h_wav.copy_(torch.from_numpy(wav_np.reshape(n, -1)))
        self._h_hot.copy_(torch.from_numpy(hot_np))
        self._d_wav.copy_(self._h_wav, non_blocking=True)
        self._d_hot.copy_(self._h_hot, non_blocking=True)
        wd = self.args.weight_decay
        vals: Dict[str, object] = {}
        if "train_op" in names:
            lr = float(self.lr_schedule(self.global_step))
            out = self.engine.train_step(self._d_wav, self._d_hot, self.params, self.slots, self.moving, lr,
                                         self.optimizer.get("momentum") or 0.0, wd, dropout_seed=self._dropout_seed(),
                                         want_outputs=bool(names & {"outputs", "logits"}))
            self.global_step += 1
            vals["train_op"], vals["learning_rate"] = None, np.float32(lr)
        else:
            out = self.engine.forward(self._d_wav, self.params, self.moving, is_training=bool(self.is_training),
                                      onehot=self._d_hot, weight_decay=wd, dropout_seed=self._dropout_seed())
            vals["learning_rate"] = np.float32(self.lr_schedule(self.global_step))
        if "metric_counts" in names:                            # evaluation counts stay on the device until read_metric_counts()
            if self._counts is None:
                self._counts = torch.zeros(self.args.num_classes ** 2 + 2, dtype=torch.int64, device=self.engine.device)
            self.engine.eval_accumulate(out["probs"], self._d_hot, self._counts, topk=5)
            vals["metric_counts"] = None
        losses = out["losses"].cpu().numpy()                    # the D2H read synchronises the step
        vals.update(total_loss=losses[0], model_loss=losses[1], global_step=np.int64(self.global_step),
                    labels=hot_np, audio_original=wav_np)
        if "outputs" in names or "logits" in names:
            vals["outputs"], vals["logits"] = out["probs"].cpu().numpy(), out["logits"].cpu().numpy()
        if "audio" in names:
            feat = self.engine.mfcc(self._d_wav)
            vals["audio"] = feat.unsqueeze(-1).cpu().numpy()
        if "ranges" in names:
            vals["ranges"] = None       # dead `fc2` head: its forward is skipped
        self._last = vals
        return vals

    @staticmethod
    def add_arguments(parser):
        parser.add_argument("--label_smoothing", default=0.0, type=float)


def _tc_flags(parser):
    parser.add_argument("--weight_decay", default=0.0001, type=float)
    parser.add_argument("--dropout_keep_prob", default=0.5, type=float)
    parser.add_argument("--width_multiplier", default=1.0, type=float)


class TCResNet8Model(AudioNetModel):
    engine_model = "TCResNet8"
    add_arguments = staticmethod(_tc_flags)


class TCResNet14Model(AudioNetModel):
    engine_model = "TCResNet14"
    add_arguments = staticmethod(_tc_flags)


def _baseline(name, flags):
    def add_arguments(parser):
        for flag, default in flags:
            if isinstance(default, str):
                parser.add_argument(flag, default=default, type=str)
            else:
                parser.add_argument(flag, default=default, type=float)
    return type(name, (AudioNetModel,), {"add_arguments": staticmethod(add_arguments), "__doc__":
                "Baseline of the reference registered for CLI compatibility; not on the accelerated path."})


_TC = [("--weight_decay", 0.0001), ("--dropout_keep_prob", 0.5), ("--width_multiplier", 1.0)]
KWSModel = _baseline("KWSModel", [("--architecture", "conv")])
Res8Model = _baseline("Res8Model", [("--weight_decay", 0.00001)])
Res8NarrowModel = _baseline("Res8NarrowModel", [("--weight_decay", 0.00001)])
Res15Model = _baseline("Res15Model", [("--weight_decay", 0.00001)])
Res15NarrowModel = _baseline("Res15NarrowModel", [("--weight_decay", 0.00001)])
DSCNNSModel = _baseline("DSCNNSModel", [("--weight_decay", 0.0)])
DSCNNMModel = _baseline("DSCNNMModel", [("--weight_decay", 0.0)])
DSCNNLModel = _baseline("DSCNNLModel", [("--weight_decay", 0.0)])
ResNet2D8Model = _baseline("ResNet2D8Model", _TC)
ResNet2D8PoolModel = _baseline("ResNet2D8PoolModel", _TC)
