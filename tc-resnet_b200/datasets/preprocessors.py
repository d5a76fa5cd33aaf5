"""Audio preprocessors with the reference's class names and call convention
(datasets/preprocessors.py:10-62, 162-203), backed by the fused CUDA front-end.

In the reference `preprocess()` appends STFT/mel/log/DCT nodes to the TF graph.  Here it records the
front-end geometry on the object (the model hands it to the engine configuration) and returns the `Node`
that stands for the feature tensor [N, T, F, 1].  `run()` is the eager form: wav tensor -> features on GPU.
The deploy (`for_deploy=True`) branch of the reference is TFLite-only and not on the accelerated path.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

from ..runtime import Node


class PreprocessorBase(ABC):
    def __init__(self, scope: str, preprocessed_node_name: str):
        self._scope = scope
        self._preprocessed_node_name = preprocessed_node_name
        self._input_node = None
        self._preprocessed_node = None

    @abstractmethod
    def preprocess(self, inputs, *args, **kwargs):
        raise NotImplementedError

    @property
    def input_node(self):
        return self._input_node

    @property
    def preprocessed_node(self):
        return self._preprocessed_node


class NoOpPreprocessor(PreprocessorBase):
    method = "no_preprocessing"

    def preprocess(self, inputs, *args, **kwargs):
        self._input_node = inputs
        self._preprocessed_node = Node(self._preprocessed_node_name, getattr(inputs, "shape", None))
        return self._preprocessed_node


class AudioPreprocessorBase(PreprocessorBase):
    method = None          # engine `preprocess_method`
    feature_key = None     # which flag gives the feature count

    def preprocess(self, inputs, window_size_samples, window_stride_samples, for_deploy=False, **kwargs):
        if for_deploy:
            raise NotImplementedError("the TFLite deploy front-end (contrib_audio.mfcc) is outside the accelerated path")
        self._input_node = inputs
        self.window_size_samples = int(window_size_samples)
        self.window_stride_samples = int(window_stride_samples)
        self.options = {k: kwargs[k] for k in ("num_mel_bins", "sample_rate", "lower_edge_hertz", "upper_edge_hertz",
                                               "num_mfccs") if k in kwargs}
        clip = inputs.shape[1] if getattr(inputs, "shape", None) else None
        frames = None if clip is None else 1 + (clip - self.window_size_samples) // self.window_stride_samples
        feats = self.options.get(self.feature_key)
        self._preprocessed_node = Node(self._preprocessed_node_name, [None, frames, feats, 1])
        return self._preprocessed_node

    def run(self, engine, wav):
        """Eager front-end on the engine's device: wav [N, L(,1)] CUDA tensor -> [N, T, F, 1]."""
        return engine.mfcc(wav).unsqueeze(-1)


class LogMelSpectrogramPreprocessor(AudioPreprocessorBase):
    method = "log_mel_spectrogram"
    feature_key = "num_mel_bins"


class MFCCPreprocessor(AudioPreprocessorBase):
    method = "mfcc"
    feature_key = "num_mfccs"
