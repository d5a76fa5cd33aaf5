"""Model ABC + the common model flags of the reference's TFModel (factory/base.py:11-67)."""
from __future__ import annotations

from abc import ABC, abstractmethod

from ..datasets.preprocessor_factory import _available_preprocessors


class TFModel(ABC):
    @staticmethod
    def add_arguments(parser):
        g = parser.add_argument_group("(CNNModel) Arguments")
        g.add_argument("--num_classes", type=int, default=None)
        g.add_argument("--checkpoint_path", default="", type=str)
        g.add_argument("--input_batch_size", type=int, default=1)
        g.add_argument("--output_name", type=str, required=True)
        g.add_argument("--preprocess_method", required=True, type=str, choices=list(_available_preprocessors))
        g.add_argument("--ignore_missing_vars", dest="ignore_missing_vars", action="store_true")
        g.add_argument("--no-ignore_missing_vars", dest="ignore_missing_vars", action="store_false")
        g.set_defaults(ignore_missing_vars=False)
        g.add_argument("--checkpoint_exclude_scopes", default="", type=str)
        g.add_argument("--checkpoint_include_scopes", default="", type=str)
        g.add_argument("--weight_decay", default=1e-4, type=float)

    @abstractmethod
    def build_deployable_model(self, *args, **kwargs):
        ...

    @abstractmethod
    def preprocess_input(self):
        ...

    @abstractmethod
    def build_output(self):
        ...

    @property
    @abstractmethod
    def audio(self):
        ...

    @property
    @abstractmethod
    def audio_original(self):
        ...

    @property
    @abstractmethod
    def total_loss(self):
        ...

    @property
    @abstractmethod
    def model_loss(self):
        ...
