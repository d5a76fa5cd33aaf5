"""Small helpers with the names the reference's common/utils.py exposes to the driver code
(positive_int :62, get_logger :69, Timer :118-139 — re-done with perf_counter because time.clock() is gone —
update_train_dir :22-59)."""
from __future__ import annotations

import argparse
import getpass
import logging
import time
from contextlib import contextmanager
from datetime import datetime

LOG_FORMAT = "[%(asctime)s] [%(name)s] %(message)s"


def positive_int(value):
    iv = int(value)
    if iv <= 0:
        raise argparse.ArgumentTypeError(f"{value} is an invalid positive int value")
    return iv


def get_logger(name=None, level=logging.INFO):
    logger = logging.getLogger(name or "tcresnet_b200")
    if not logger.handlers:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter(LOG_FORMAT))
        logger.addHandler(handler)
        logger.setLevel(level)
        logger.propagate = False
    return logger


class Timer:
    def __init__(self, log=None):
        self.log = log

    @contextmanager
    def __call__(self, name, log_fn=None):
        t0 = time.perf_counter()
        yield
        msg = f"{name}: {(time.perf_counter() - t0) * 1e3:.1f} ms"
        (log_fn or (self.log.info if self.log else print))(msg)


def update_train_dir(args):
    """%DATE% / %USER% / %KEY% templating of --train_dir."""
    if getattr(args, "train_dir", None):
        d = args.train_dir.replace("%DATE%", datetime.now().strftime("%y%m%d%H%M%S")).replace("%USER%", getpass.getuser())
        args.train_dir = d.replace("%KEY%", getattr(args, "model", "model"))
    return args
