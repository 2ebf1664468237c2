"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

NumPy/SciPy restatement of the operator builders that produce the *inputs* of
the hot path (operator construction itself is out of scope for the kernels,
SURVEY.md section 2 rows 11-13).  Used to generate synthetic operators for
parity tests and the bench.
"""

from __future__ import annotations

from math import comb

import numpy as np
from scipy.linalg import expm


def annihilation_operator(cutoff: int) -> np.ndarray:
    """_math/ops.py:230-239."""
    return np.diag(np.sqrt(np.arange(1, cutoff, dtype=np.complex128)), 1)


def creation_operator(cutoff: int) -> np.ndarray:
    """_math/ops.py:242-251."""
    return np.conjugate(annihilation_operator(cutoff)).T


def number_operator(cutoff: int) -> np.ndarray:
    """_math/ops.py:254-263."""
    return creation_operator(cutoff) @ annihilation_operator(cutoff)


def squeezing_operator(cutoff: int, zeta: complex) -> np.ndarray:
    """_math/ops.py:266-281."""
    c, d = creation_operator(cutoff), annihilation_operator(cutoff)
    return expm(0.5 * (np.conj(zeta) * (d @ d) - zeta * (c @ c)))


def displacement_operator(cutoff: int, alpha: complex) -> np.ndarray:
    """_math/ops.py:284-298."""
    c, d = creation_operator(cutoff), annihilation_operator(cutoff)
    return expm(alpha * c - np.conj(alpha) * d)


def phase_operator(cutoff: int, theta: float) -> np.ndarray:
    """_math/ops.py:301-333."""
    return np.diag(np.exp(1j * np.arange(cutoff) * theta))


def beam_splitter(d0: int, d1: int, eta: float) -> np.ndarray:
    """
    operation/composite_operation.py:292-299 (NonPolarizingBeamSplitter).  The
    reference's variable names are swapped (``a = creation_operator``), the
    arithmetic below is what it computes.
    """
    a = creation_operator(d0)
    a_dagger = annihilation_operator(d0)
    b = creation_operator(d1)
    b_dagger = annihilation_operator(d1)
    return expm(1j * eta * (np.kron(a_dagger, b) + np.kron(a, b_dagger)))


def hadamard_operator() -> np.ndarray:
    return np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2)


def x_operator() -> np.ndarray:
    return np.array([[0, 1], [1, 0]], dtype=np.complex128)


def rx_operator(theta: float) -> np.ndarray:
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)


def loss_channel(cutoff: int, gamma: float) -> np.ndarray:
    """
    Photon-loss (amplitude damping) channel, not shipped by the reference
    (SURVEY.md section 8d): K_k = sum_n sqrt(C(n,k) (1-g)^(n-k) g^k) |n-k><n|,
    k = 0..cutoff-1.  Returns (cutoff, cutoff, cutoff); sum_k K^dagger K = I.
    """
    ks = np.zeros((cutoff, cutoff, cutoff), dtype=np.complex128)
    for k in range(cutoff):
        for n in range(k, cutoff):
            ks[k, n - k, n] = np.sqrt(comb(n, k) * (1 - gamma) ** (n - k) * gamma**k)
    return ks


def random_unitary(d: int, seed: int) -> np.ndarray:
    """Haar-ish dense unitary via QR of a complex Gaussian (worst-case dense op)."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
    q, r = np.linalg.qr(z)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def random_state(n: int, seed: int, dtype=np.complex128) -> np.ndarray:
    """Normalised complex Gaussian state vector shaped (n, 1)."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    v /= np.linalg.norm(v)
    return v.reshape(n, 1).astype(dtype)


def random_density(n: int, seed: int, rank: int = 4, dtype=np.complex128) -> np.ndarray:
    """rho = A A^dagger / Tr with A an (n, rank) complex Gaussian (PSD, trace 1)."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, rank)) + 1j * rng.standard_normal((n, rank))
    rho = a @ np.conj(a.T)
    return (rho / np.trace(rho)).astype(dtype)
