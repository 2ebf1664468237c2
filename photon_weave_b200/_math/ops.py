"""
Operator builders (reference: photon_weave/_math/ops.py:17-512).

Operators are tiny (t x t, t = product of the target dimensions) and are built once on
the host in complex128 with NumPy/SciPy, then cached in HBM by ``Operation`` -- the
reference re-runs ``expm``/``kron`` on every ``Operation.operator`` access
(operation/operation.py:292-294).  Everything that touches a *state* is a CUDA kernel;
the only state-sized helpers here (``apply_kraus``, ``num_quanta_*``) go through the
device facade.
"""

from __future__ import annotations

from typing import Any, List, Sequence, Union

import numpy as np
from scipy.linalg import expm as _expm

from ..core import device as dev

_C = np.complex128
_S2 = 1 / np.sqrt(2)


def jitted_exp(matrix: Any) -> np.ndarray:
    return _expm(np.asarray(matrix, dtype=_C))


def identity_operator() -> np.ndarray:
    return np.eye(2)


def hadamard_operator() -> np.ndarray:
    return np.array([[1, 1], [1, -1]]) * _S2


def x_operator() -> np.ndarray:
    return np.array([[0, 1], [1, 0]])


def y_operator() -> np.ndarray:
    return np.array([[0, -1j], [1j, 0]])


def z_operator() -> np.ndarray:
    return np.array([[1, 0], [0, -1]])


def s_operator() -> np.ndarray:
    return np.array([[1, 0], [0, 1j]])


def t_operator() -> np.ndarray:
    return np.array([[1, 0], [0, np.exp(1j * np.pi / 4)]])


def controlled_not_operator() -> np.ndarray:
    return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])


def controlled_z_operator() -> np.ndarray:
    return np.diag([1, 1, 1, -1])


def swap_operator() -> np.ndarray:
    return np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])


def sx_operator() -> np.ndarray:
    return np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]]) / 2


def controlled_swap_operator() -> np.ndarray:
    m = np.eye(8)
    m[[5, 6]] = m[[6, 5]]
    return m


def rx_operator(theta: float) -> np.ndarray:
    c, s = np.cos(theta / 2), 1j * np.sin(-theta / 2)
    return np.array([[c, s], [s, c]])


def ry_operator(theta: float) -> np.ndarray:
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return np.array([[c, -s], [s, c]])


def rz_operator(theta: float) -> np.ndarray:
    return np.array([[np.exp(-1j * theta / 2), 0], [0, np.exp(1j * theta / 2)]])


def u3_operator(phi: float, theta: float, omega: float) -> np.ndarray:
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return np.array([[c, -np.exp(1j * omega) * s],
                     [np.exp(1j * phi) * s, np.exp(1j * (phi + omega)) * c]])


def annihilation_operator(cutoff: int) -> np.ndarray:
    return np.diag(np.sqrt(np.arange(1, cutoff, dtype=_C)), 1)


def creation_operator(cutoff: int) -> np.ndarray:
    return np.conjugate(annihilation_operator(cutoff)).T


def number_operator(cutoff: int) -> np.ndarray:
    return creation_operator(cutoff) @ annihilation_operator(cutoff)


def squeezing_operator(cutoff: int, zeta: complex) -> np.ndarray:
    a, ad = annihilation_operator(cutoff), creation_operator(cutoff)
    return _expm(0.5 * (np.conj(zeta) * (a @ a) - zeta * (ad @ ad)))


def displacement_operator(cutoff: int, alpha: complex) -> np.ndarray:
    a, ad = annihilation_operator(cutoff), creation_operator(cutoff)
    return _expm(alpha * ad - np.conj(alpha) * a)


def phase_operator(cutoff: int, theta: float) -> np.ndarray:
    return np.diag(np.exp(1j * np.arange(cutoff) * theta))


def _stack_kraus_operators(operators: Union[Any, Sequence[Any]]) -> np.ndarray:
    """_math/ops.py:343-366."""
    ops = np.asarray([dev.to_numpy(o) for o in operators]) if isinstance(operators, (list, tuple)) \
        else dev.to_numpy(operators)
    if ops.ndim == 2:
        ops = ops[None, ...]
    if ops.ndim != 3:
        raise ValueError(f"Kraus operators must have shape (k, d, d) or (d, d); got {ops.shape}")
    return ops


def apply_kraus(density_matrix: Any, kraus_operators: Union[Any, List[Any]]) -> Any:
    """sum_k K rho K^dagger on an un-embedded state (_math/ops.py:371-403): the Kraus kernel
    with a single subsystem of dimension d."""
    from ..core import adapters

    operators = _stack_kraus_operators(kraus_operators)
    shp = tuple(density_matrix.shape)
    if shp != operators.shape[-2:]:
        raise ValueError(f"Density matrix shape {shp} does not match Kraus operator dims {operators.shape[-2:]}")
    return adapters.apply_kraus_matrix((shp[0],), (0,), density_matrix, operators)


def kraus_identity_check(operators: Union[Any, List[Any]], tol: float = 1e-6) -> bool:
    """sum_k K^dagger K <= I within tol (_math/ops.py:406-431).  Operator-sized host check."""
    ops = _stack_kraus_operators(operators).astype(_C)
    kraus_sum = np.einsum("kia,kib->ab", np.conj(ops), ops)
    eigvals = np.linalg.eigvalsh(np.eye(kraus_sum.shape[0], dtype=_C) - kraus_sum)
    return bool(np.all(eigvals >= -tol))


def num_quanta_vector(vector: Any) -> int:
    """Highest row with a non-zero entry (_math/ops.py:472-486)."""
    shp = dev.shape(vector)
    levels, _, _ = dev.occupancy(vector, (shp[0],), len(shp) == 2 and shp[1] != 1)
    return int(levels[0])


def num_quanta_matrix(matrix: Any) -> int:
    """Highest row/column with a non-zero entry (_math/ops.py:489-512).  Only ever called on
    reduced density matrices (positive semi-definite), where that is the highest non-zero
    diagonal entry."""
    shp = dev.shape(matrix)
    levels, _, _ = dev.occupancy(matrix, (shp[0],), True)
    return int(levels[0])
