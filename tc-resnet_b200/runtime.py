"""Minimal stand-ins for the two TensorFlow objects the reference's driver code touches: graph tensors
(`Node`) and `tf.Session` (`Session`).

The reference builds a TF graph once and calls `session.run(fetch_dict, feed_dict)` per step
(helper/trainer.py:312-321, helper/base.py:95-104).  Here a `Node` is a named handle; `Session.run`
asks the bound model to execute ONE step of the CUDA engine for the requested names — a training step when a
train op is among them, a forward pass otherwise — and returns NumPy values in the structure that was passed
in (dict / list / single node), exactly what the trainer and evaluator loops consume.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, Optional


class Node:
    __slots__ = ("name", "shape", "owner")

    def __init__(self, name: str, shape=None, owner=None):
        self.name, self.shape, self.owner = name, shape, owner

    def __repr__(self):
        return f"<Node {self.name} shape={self.shape}>"


class OutOfRangeError(Exception):
    """End of a non-repeating dataset (tf.errors.OutOfRangeError in helper/base.py:117)."""


class InvalidArgumentError(Exception):
    """A malformed sample; the reference skips the step (helper/trainer.py:430, helper/base.py:120)."""


class Session:
    def __init__(self, config=None):
        self.config = config
        self._executor = None

    def bind(self, executor):
        self._executor = executor

    @staticmethod
    def _walk(fetches) -> Iterable[Node]:
        if isinstance(fetches, Node):
            yield fetches
        elif isinstance(fetches, dict):
            for v in fetches.values():
                yield from Session._walk(v)
        elif isinstance(fetches, (list, tuple)):
            for v in fetches:
                yield from Session._walk(v)
        elif fetches is None:
            return
        else:
            raise TypeError(f"cannot fetch {type(fetches).__name__}")

    def run(self, fetches, feed_dict: Optional[Dict[Any, Any]] = None):
        nodes = list(self._walk(fetches))
        names = {n.name for n in nodes}
        values: Dict[str, Any] = {}
        runnable = names - {"noop", "init"}
        if runnable:
            if self._executor is None:
                raise RuntimeError("session.run before model.build(): nothing is bound to this session")
            values = self._executor.execute(runnable, feed_dict or {})
        values.setdefault("noop", None)
        values.setdefault("init", None)

        def rebuild(f):
            if isinstance(f, Node):
                return values[f.name]
            if isinstance(f, dict):
                return {k: rebuild(v) for k, v in f.items()}
            if isinstance(f, (list, tuple)):
                return type(f)(rebuild(v) for v in f)
            return None
        return rebuild(fetches)


def global_variables_initializer():
    return Node("init")


def local_variables_initializer():
    return Node("init")


def no_op():
    return Node("noop")
