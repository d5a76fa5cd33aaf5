"""Constants the reference keeps in const.py:4-9.  TF_SESSION_CONFIG has no meaning here (kept as a name so
`Session(config=const.TF_SESSION_CONFIG)` call sites read the same); the single-device pin `device_count={"GPU": 1}`
is replaced by one process per GPU under torchrun."""
TF_SESSION_CONFIG = None
NULL_CLASS_LABEL = "__null__"
BACKGROUND_NOISE_DIR_NAME = "_background_noise_"
