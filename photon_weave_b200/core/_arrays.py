"""
Array plumbing for the seam: device memory, streams and zero-copy DLPack import
come from PyTorch; every arithmetic step is a libpwb200 kernel.

Accepted inputs for states / operators:
* ``torch.Tensor`` on a CUDA device  (native),
* any object exporting ``__dlpack__`` on a CUDA device (JAX, CuPy ...), imported
  zero-copy with ``torch.from_dlpack``,
* host data (NumPy arrays, lists): copied to the current CUDA device; results of
  such calls are returned as NumPy arrays so the call is symmetric.
"""

from __future__ import annotations

from typing import Any, Tuple

import numpy as np
import torch

from .. import _lib

_COMPLEX = {torch.complex64: _lib.PW_C64, torch.complex128: _lib.PW_C128}
_REAL_OF = {torch.complex64: torch.float32, torch.complex128: torch.float64}


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "photon_weave_b200 needs a CUDA device (B200, sm_100a); there is no CPU path"
        )


def device() -> str:
    """Where new arrays are allocated (the current CUDA device)."""
    return "cuda"


try:  # the raw handle of the current stream without building a torch.cuda.Stream object (~1 us)
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # pragma: no cover
    _raw_stream = None


_get_device = getattr(torch._C, "_cuda_getDevice", None)   # torch.cuda.current_device() minus its lazy-init check


def stream_ptr() -> int:
    if _raw_stream is not None:
        return _raw_stream(_get_device() if _get_device is not None else torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def is_host(x: Any) -> bool:
    if isinstance(x, torch.Tensor):
        return not x.is_cuda
    if hasattr(x, "__dlpack_device__"):
        try:
            return int(x.__dlpack_device__()[0]) not in (2, 13)  # kDLCUDA, kDLCUDAManaged
        except Exception:
            return True
    return True


def to_device(x: Any) -> torch.Tensor:
    """Return a CUDA tensor viewing (device input) or holding a copy of (host input) x."""
    if isinstance(x, torch.Tensor):
        return x if x.is_cuda else x.cuda()
    if hasattr(x, "__dlpack__") and not is_host(x):
        return torch.from_dlpack(x)
    return torch.as_tensor(np.asarray(x)).cuda()


def complex_dtype_for(*tensors: torch.Tensor) -> torch.dtype:
    """JAX-style promotion restricted to the two complex types the kernels compute in
    (x64 is always on in the reference: photon_weave/_math/ops.py:11)."""
    dt = tensors[0].dtype
    for t in tensors[1:]:
        dt = torch.promote_types(dt, t.dtype)
    if dt == torch.complex64:
        return torch.complex64
    if dt in (torch.float32, torch.float16, torch.bfloat16):
        return torch.complex64
    return torch.complex128


def as_complex(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    if x.dtype != dtype:
        x = x.to(dtype)
    return x.contiguous()


def code(dtype: torch.dtype) -> int:
    return _COMPLEX[dtype]


def real_dtype(dtype: torch.dtype) -> torch.dtype:
    return _REAL_OF[dtype]


def back(x: torch.Tensor, host: bool):
    """Return the result in the caller's world (NumPy for host callers)."""
    return x.cpu().numpy() if host else x


def key_words(key: Any) -> Tuple[int, int]:
    """A PRNG key is two uint32 words (jax.random.PRNGKey layout)."""
    if isinstance(key, torch.Tensor):
        key = key.detach().cpu().numpy()
    k = np.asarray(key).astype(np.uint32).reshape(-1)
    if k.size != 2:
        raise ValueError(f"PRNG key must hold two uint32 words, got shape {np.shape(key)}")
    return int(k[0]), int(k[1])
