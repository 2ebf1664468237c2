"""NumPy-backed subset of jax.numpy (TEST INFRASTRUCTURE ONLY, see jax/__init__.py)."""

import numpy as _np
import torch as _torch

pi = _np.pi
complex128 = _np.complex128
complex64 = _np.complex64
float64 = _np.float64
float32 = _np.float32
int32 = _np.int32
int64 = _np.int64
newaxis = None


class _At:
    def __init__(self, arr):
        self.arr = arr
        self.idx = None

    def __getitem__(self, idx):
        self.idx = idx
        return self

    def set(self, value):
        out = _np.array(self.arr)
        out[self.idx] = value
        return _wrap(out)

    def add(self, value):
        out = _np.array(self.arr)
        out[self.idx] += value
        return _wrap(out)


class _Meta(type(_np.ndarray)):
    def __instancecheck__(cls, obj):
        return isinstance(obj, (_np.ndarray, _torch.Tensor))


class ndarray(_np.ndarray, metaclass=_Meta):
    @property
    def at(self):
        return _At(self)


def _host(x):
    if isinstance(x, _torch.Tensor):
        return x.detach().cpu().numpy()
    if isinstance(x, (list, tuple)):
        return type(x)(_host(v) for v in x)
    return x


def _wrap(x):
    if isinstance(x, _np.ndarray):
        return x.view(ndarray)
    return x


def _lift(fn):
    def wrapped(*args, **kwargs):
        args = [_host(a) for a in args]
        kwargs = {k: _host(v) for k, v in kwargs.items()}
        return _wrap(fn(*args, **kwargs))

    wrapped.__name__ = getattr(fn, "__name__", "fn")
    return wrapped


for _name in ["array", "asarray", "allclose", "sqrt", "zeros", "conj", "conjugate", "eye", "kron", "sin",
              "cos", "isfinite", "stack", "exp", "isclose", "diag", "arange", "trace", "real", "imag",
              "ones_like", "zeros_like", "matmul", "copy", "vstack", "hstack", "transpose", "reshape",
              "clip", "all", "any", "array_equal", "ones", "abs", "sum", "prod", "outer", "dot", "angle",
              "identity", "linspace", "max", "min", "argmax", "where", "nonzero", "concatenate", "log",
              "power", "cumsum", "tensordot", "einsum", "squeeze", "expand_dims", "isnan", "pad", "take",
              "mean", "std", "var", "round", "floor", "ceil", "log2", "log10", "tan", "arctan2", "sign", "mod",
              "maximum", "minimum", "argmin", "sort", "unique", "flip", "roll", "tile", "repeat", "full", "empty"]:
    globals()[_name] = _lift(getattr(_np, _name))


class linalg:
    norm = staticmethod(_lift(_np.linalg.norm))
    inv = staticmethod(_lift(_np.linalg.inv))
    eigvals = staticmethod(_lift(_np.linalg.eigvals))
    eigh = staticmethod(_lift(_np.linalg.eigh))
    eigvalsh = staticmethod(_lift(_np.linalg.eigvalsh))
    det = staticmethod(_lift(_np.linalg.det))
