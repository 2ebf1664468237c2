"""
Array facade for the state containers (Fock / Polarization / CustomState / Envelope /
CompositeEnvelope): every array the containers hold lives in HBM as a ``torch.Tensor``
and every operation on it is a libpwb200 kernel reached through ``core.ffi``.

The containers never touch array contents directly -- they call the functions below
and ``core.adapters`` -- so their host logic (routing, ordering, bookkeeping) is
independent of where the array lives.  Functions marked [sync] copy a few scalars
back to the host because the caller branches on them, exactly where the reference
calls ``int(...)`` / ``bool(...)`` on a JAX array.
"""

from __future__ import annotations

from typing import Any, List, Sequence, Tuple

import numpy as np
import torch

from . import _arrays as A
from . import ffi

DTYPE = torch.complex128  # the reference computes in complex128 throughout (_math/ops.py:11)


def is_array(x: Any) -> bool:
    return isinstance(x, torch.Tensor)


def asarray(x: Any) -> torch.Tensor:
    """Host data / device array -> contiguous complex128 CUDA tensor."""
    A.require_cuda()
    t = A.to_device(x)
    return A.as_complex(t, DTYPE)


def pin_operator(op: torch.Tensor) -> torch.Tensor:
    """Tell libpwb200 that a cached device operator never changes (pw_op_pin): its structure
    analysis is then kept across applies.  No-op for host tensors (the oracle-backed test shim)."""
    if isinstance(op, torch.Tensor) and op.is_cuda:
        ffi.pin_operator(op)
    return op


def adopt(value: Any) -> Any:
    """What a container stores when the user assigns ``.state``: labels pass through, host
    arrays (NumPy / JAX / lists of rows) move into HBM as complex128."""
    if isinstance(value, torch.Tensor):
        return value if value.is_complex() and value.device.type == A.device() else asarray(value)
    if hasattr(value, "shape") and hasattr(value, "dtype") and len(getattr(value, "shape")) >= 1:
        return asarray(value)
    return value


def to_numpy(x: Any) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def shape(x: Any) -> Tuple[int, ...]:
    return tuple(x.shape)


def nbytes(x: Any) -> int:
    return int(x.numel() * x.element_size())


def basis_vector(dim: int, index: int) -> torch.Tensor:
    """|index> in a dim-dimensional space, shape (dim, 1) (state_transform.py:46-52)."""
    A.require_cuda()
    v = torch.zeros((int(dim), 1), dtype=DTYPE, device=A.device())
    v[int(index), 0] = 1
    return v


def scalar_one() -> torch.Tensor:
    """The empty product state [[1]] (composite_envelope.py:103)."""
    A.require_cuda()
    return torch.ones((1, 1), dtype=DTYPE, device=A.device())


def kron(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return ffi.kron(a.contiguous(), b.contiguous())


def expand(psi: torch.Tensor) -> torch.Tensor:
    """psi -> psi psi^dagger."""
    return ffi.expand_vec(psi.contiguous())


def contract(rho: torch.Tensor, tol: float = 1e-6):
    """[sync] (psi or None, purity, ok): Matrix -> Vector when |Tr rho^2 - 1| < tol."""
    psi, info, flag = ffi.contract_mat(rho.contiguous(), tol)
    ok = bool(int(flag.item()))
    return (psi if ok else None), float(info[0].item()), ok


def find_label(psi: torch.Tensor) -> Tuple[int, int]:
    """[sync] (number of entries == 1+0j, index of the first one)."""
    out = ffi.find_label(psi.contiguous()).tolist()
    return int(out[0]), int(out[1])


def resize(x: torch.Tensor, dims: Sequence[int], axis: int, new_dim: int, is_matrix: bool) -> torch.Tensor:
    return ffi.resize_axis(x.contiguous(), list(dims), axis, new_dim, is_matrix)


def permute(x: torch.Tensor, dims: Sequence[int], perm: Sequence[int], is_matrix: bool) -> torch.Tensor:
    """Reorder the subsystems of a product state: new axis k = old axis perm[k]."""
    dims = list(dims)
    perm = list(perm)
    n = len(dims)
    if is_matrix:
        out = ffi.permute(x.contiguous(), dims + dims, perm + [n + p for p in perm])
        N = int(np.prod(dims))
        return out.reshape(N, N)
    return ffi.permute(x.contiguous(), dims, perm).reshape(-1, 1)


def occupancy(x: torch.Tensor, dims: Sequence[int], is_matrix: bool):
    """[sync] (levels, total, max_abs2): highest occupied basis index of every subsystem in
    ONE pass, plus sum |psi|^2 (or Tr rho) and max |x|^2."""
    levels, stats = ffi.occupancy(x.contiguous(), list(dims), is_matrix)
    s = stats.tolist()
    return [int(v) for v in levels.tolist()], float(s[0]), float(s[1])


def norm_stats(x: torch.Tensor) -> Tuple[float, float]:
    """[sync] (sum |x|^2, max |x|^2) over all entries."""
    s = ffi.norm_stats(x.contiguous()).tolist()
    return float(s[0]), float(s[1])


def draw(probabilities: Any, key: Any) -> int:
    """[sync] jax.random.choice(key, arange(len(p)), p=p) for a host vector of weights:
    uploaded once, sampled by ``pw_draw`` (same threefry bits / cumsum / searchsorted as the
    measurement kernels)."""
    A.require_cuda()
    p = torch.as_tensor(np.asarray(probabilities, dtype=np.float64)).to(A.device())
    return int(ffi.draw(p.contiguous(), key, DTYPE).item())


def scale(x: torch.Tensor, factor: float) -> torch.Tensor:
    """x *= factor (in place on a contiguous tensor)."""
    x = x.contiguous()
    f = torch.full((1,), float(factor), dtype=torch.float64, device=x.device)
    return ffi.scale_(x, f)


def divide(x: torch.Tensor, divisor: float) -> torch.Tensor:
    """x /= divisor (in place, true division: ``state / norm`` with the reference's rounding)."""
    x = x.contiguous()
    d = torch.full((1,), float(divisor), dtype=torch.float64, device=x.device)
    return ffi.divide_(x, d)


def entries(x: torch.Tensor) -> List[complex]:
    """[sync] small arrays only: the entries as Python complex numbers."""
    return [complex(v) for v in to_numpy(x).reshape(-1)]
