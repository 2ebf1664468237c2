"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

NumPy restatement of ``photon_weave/core/kernels.py`` ("ordered" primitives:
target subsystems already lead the tensor order).  Each function cites the
reference lines it follows.  Contractions are written with ``np.einsum`` /
``np.tensordot`` using the same index strings as the reference; a tensor
contraction is order-insensitive to ~1e-16 relative in complex128, well inside
the 1e-12 parity tolerance.

Sampling (``jax.random.choice``) is restated in ``rng_threefry.choice``.
"""

from __future__ import annotations

from math import prod
from typing import Sequence, Tuple

import numpy as np

from . import rng_threefry as rng


def _real_dtype(x: np.ndarray):
    return np.float32 if x.dtype == np.complex64 else np.float64


def measurement_probs_vector(state_dims, target_count, product_state):
    """core/kernels.py:28-41  probs = sum_rest |psi|^2, then / sum."""
    t = prod(state_dims[:target_count])
    r = prod(state_dims[target_count:])
    ps = np.asarray(product_state).reshape((t, r, 1))
    probs = np.einsum("tcr->t", np.abs(ps) ** 2)
    probs = probs / np.sum(probs)
    return probs, ps, r


def measurement_probs_matrix(state_dims, target_count, product_state):
    """core/kernels.py:44-59  probs = real(diag(Tr_rest rho)) / sum."""
    t = prod(state_dims[:target_count])
    r = prod(state_dims[target_count:])
    ps = np.asarray(product_state).reshape((t, r, t, r))
    ps = np.transpose(ps, (0, 2, 1, 3))
    reduced = np.einsum("ijrr->ij", ps)
    probs = np.real(np.diag(reduced))
    probs = probs / np.sum(probs)
    return probs, ps, r


def apply_op_vector_ordered(state_dims, target_count, product_state, operator):
    """core/kernels.py:62-98  'ab,bcn->acn' on (t, r, 1)."""
    t = prod(state_dims[:target_count])
    r = prod(state_dims[target_count:])
    product_state = np.asarray(product_state)
    dt = np.result_type(product_state.dtype, np.asarray(operator).dtype)
    ps = product_state.reshape((t, r)).astype(dt, copy=False)
    op = np.asarray(operator).reshape((t, t)).astype(dt, copy=False)
    return (op @ ps).reshape((t * r, 1))


def apply_op_matrix_ordered(state_dims, target_count, product_state, operator):
    """core/kernels.py:101-145  rho' = (O x I) rho (O x I)^dagger."""
    t = prod(state_dims[:target_count])
    r = prod(state_dims[target_count:])
    product_state = np.asarray(product_state)
    dt = np.result_type(product_state.dtype, np.asarray(operator).dtype)
    ps = product_state.reshape((t, r, t, r)).astype(dt, copy=False)
    ps = np.transpose(ps, (0, 2, 1, 3)).reshape((t, t, r * r))
    op = np.asarray(operator).reshape((t, t)).astype(dt, copy=False)
    tmp = np.einsum("ab,bcn->acn", op, ps, optimize=True)
    out = np.einsum("acn,bc->abn", tmp, np.conj(op), optimize=True)
    out = out.reshape((t, t, r, r))
    return np.transpose(out, (0, 2, 1, 3)).reshape((t * r, t * r))


def apply_kraus_matrix_ordered(state_dims, target_count, product_state, operators):
    """core/kernels.py:148-186  sum_k of apply_op_matrix_ordered (vmap + sum)."""
    operators = np.asarray(operators)
    acc = None
    for k in range(operators.shape[0]):
        term = apply_op_matrix_ordered(
            state_dims, target_count, product_state, operators[k]
        )
        acc = term if acc is None else acc + term
    return acc


def measure_vector_ordered(state_dims, target_count, product_state, key):
    """core/kernels.py:189-206; post = psi[o] / sqrt(p_o + 1e-18); key unchanged."""
    probs, ps, r = measurement_probs_vector(state_dims, target_count, product_state)
    outcome = rng.choice(key, ps.shape[0], probs)
    post = ps[outcome] / np.sqrt(probs[outcome] + probs.dtype.type(1e-18))
    return outcome, post.reshape((r, 1)), key


def measure_matrix_ordered(state_dims, target_count, product_state, key):
    """core/kernels.py:209-227; post = rho[o,o] / (p_o + 1e-18)."""
    probs, ps, r = measurement_probs_matrix(state_dims, target_count, product_state)
    outcome = rng.choice(key, ps.shape[0], probs)
    post = ps[outcome, outcome] / (probs[outcome] + probs.dtype.type(1e-18))
    return outcome, post.reshape((r, r)), key


def measure_vector_ordered_with_probs(state_dims, target_count, product_state, key):
    """core/kernels.py:230-245."""
    probs, ps, r = measurement_probs_vector(state_dims, target_count, product_state)
    outcome = rng.choice(key, ps.shape[0], probs)
    post = ps[outcome] / np.sqrt(probs[outcome] + probs.dtype.type(1e-18))
    return outcome, post.reshape((r, 1)), probs, key


def measure_matrix_ordered_with_probs(state_dims, target_count, product_state, key):
    """core/kernels.py:248-263."""
    probs, ps, r = measurement_probs_matrix(state_dims, target_count, product_state)
    outcome = rng.choice(key, ps.shape[0], probs)
    post = ps[outcome, outcome] / (probs[outcome] + probs.dtype.type(1e-18))
    return outcome, post.reshape((r, r)), probs, key


def measure_vector_expectation_ordered(state_dims, target_count, product_state):
    """core/kernels.py:266-283  'tr,ts,t->rs' of the normalised branches."""
    probs, ps, r = measurement_probs_vector(state_dims, target_count, product_state)
    norm_states = ps[..., 0] / np.sqrt(probs[:, None] + probs.dtype.type(1e-18))
    expected = np.einsum("tr,ts,t->rs", norm_states, np.conj(norm_states), probs)
    return probs, expected.reshape((r, r))


def trace_out_matrix_ordered(state_dims, target_count, product_state):
    """core/kernels.py:301-319  trace out the LEADING target_count axes ('iin->n')."""
    t = prod(state_dims[:target_count])
    r = prod(state_dims[target_count:])
    ps = np.asarray(product_state).reshape((t, r, t, r))
    ps = np.transpose(ps, (0, 2, 1, 3)).reshape((t, t, r * r))
    traced = np.einsum("iin->n", ps)
    return traced.reshape((r, r))


def measure_matrix_expectation_ordered(state_dims, target_count, product_state):
    """core/kernels.py:286-298."""
    probs, _, _ = measurement_probs_matrix(state_dims, target_count, product_state)
    expected = trace_out_matrix_ordered(state_dims, target_count, product_state)
    return probs, expected


def measure_povm_matrix_ordered(state_dims, target_count, operators, product_state, key):
    """core/kernels.py:322-344  E_k rho E_k^dagger for all k, p_k = real Tr, choice."""
    operators = np.asarray(operators)
    projected = np.stack(
        [
            apply_op_matrix_ordered(state_dims, target_count, product_state, operators[k])
            for k in range(operators.shape[0])
        ]
    )
    probs = np.real(np.einsum("kii->k", projected))
    probs = probs / np.sum(probs)
    outcome = rng.choice(key, operators.shape[0], probs)
    post = projected[outcome]
    post = post / (np.trace(post) + 1e-18)
    return outcome, post, key
