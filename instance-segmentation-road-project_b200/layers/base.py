"""Minimal stand-in for tf.keras.layers.Layer: the hot-path layers of the reference are
stateless (no weights), so all the interface they use is `__init__(**kwargs)`,
`__call__(inputs, **call_kwargs) -> call(...)`, `get_config()` / `from_config()` and the
by-name registry `get_custom_objects()` (/root/reference/engine/__init__.py:26-55).
"""
import ctypes

import torch

from .. import runtime as rt

_CUSTOM_OBJECTS = {}
_NAME_COUNTS = {}


def get_custom_objects():
    return _CUSTOM_OBJECTS


def register(cls):
    _CUSTOM_OBJECTS[cls.__name__] = cls
    return cls


def _snake(name):
    out = []
    for i, ch in enumerate(name):
        if ch.isupper() and i and not name[i - 1].isupper():
            out.append("_")
        out.append(ch.lower())
    return "".join(out)


class Layer:
    def __init__(self, name=None, trainable=True, dtype="float32", **kwargs):
        if kwargs:
            raise TypeError(f"{type(self).__name__}: unexpected keyword arguments {sorted(kwargs)}")
        if name is None:
            base = _snake(type(self).__name__)
            n = _NAME_COUNTS.get(base, 0)
            _NAME_COUNTS[base] = n + 1
            name = base if n == 0 else f"{base}_{n}"
        self.name = name
        self.trainable = trainable
        self.dtype = dtype

    def __call__(self, inputs, **kwargs):
        return self.call(inputs, **kwargs)

    def call(self, inputs, **kwargs):          # pragma: no cover
        raise NotImplementedError

    def get_config(self):
        return {"name": self.name, "trainable": self.trainable, "dtype": self.dtype}

    @classmethod
    def from_config(cls, config):
        return cls(**config)


def ctx_of(*tensors):
    for t in tensors:
        if isinstance(t, torch.Tensor):
            if not t.is_cuda:
                raise rt.InvalidArgumentError(
                    rt.MLP_EDLPACK, f"tensor on {t.device}: masklab_b200 has no CPU path")
            return rt.Context.get(t.device)
    return rt.Context.get()


def i32_scalar(ctx, n=1):
    return torch.empty((n,), dtype=torch.int32, device=ctx.torch_device)


def null():
    return ctypes.c_void_p(None)
