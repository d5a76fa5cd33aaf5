"""name -> preprocessor registry (datasets/preprocessor_factory.py:6-19)."""
from .preprocessors import LogMelSpectrogramPreprocessor, MFCCPreprocessor, NoOpPreprocessor

_available_preprocessors = {
    "log_mel_spectrogram": LogMelSpectrogramPreprocessor,
    "mfcc": MFCCPreprocessor,
    "no_preprocessing": NoOpPreprocessor,
}


def factory(preprocess_method, scope, preprocessed_node_name):
    try:
        cls = _available_preprocessors[preprocess_method]
    except KeyError:
        raise NotImplementedError(f"{preprocess_method}") from None
    return cls(scope, preprocessed_node_name)
