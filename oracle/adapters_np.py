"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

NumPy restatement of the array-only seam ``photon_weave/core/adapters.py``
(the drop-in boundary): same names, argument order,
return shapes and error behaviour.  ``use_contraction=True`` follows the
reference's einsum-on-the-unpermuted-tensor route (adapters.py:112-166, strings
from oracle/einsum_strings.py); ``False`` follows reorder -> ordered kernel ->
restore (adapters.py:210-296, 334-345).
"""

from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

from . import einsum_strings as es
from . import kernels_np as kernels


def _permute_dims(dims, perm):
    return [dims[i] for i in perm]


def _inverse_perm(perm):
    inv = [0] * len(perm)
    for i, p in enumerate(perm):
        inv[p] = i
    return inv


def _reorder_meta(state_dims, target_indices):
    """adapters.py:42-70."""
    n = len(state_dims)
    rest = [i for i in range(n) if i not in target_indices]
    perm = list(target_indices) + rest
    return perm, rest, _permute_dims(state_dims, perm)


def reorder_vector_front(state_dims, target_indices, product_state):
    """adapters.py:210-231."""
    n = len(state_dims)
    perm, _, reordered = _reorder_meta(tuple(state_dims), tuple(target_indices))
    if perm == list(range(n)):
        return product_state, list(state_dims), perm
    ps = np.asarray(product_state).reshape((*state_dims, 1))
    ps = np.transpose(ps, perm + [n]).reshape((-1, 1))
    return ps, reordered, perm


def restore_vector_order(state_dims, perm, product_state):
    """adapters.py:234-248."""
    if list(perm) == list(range(len(perm))):
        return product_state
    inv = _inverse_perm(perm)
    reordered = _permute_dims(state_dims, perm)
    ps = np.asarray(product_state).reshape((*reordered, 1))
    ps = np.transpose(ps, inv + [len(inv)])
    return ps.reshape((math.prod(state_dims), 1))


def reorder_matrix_front(state_dims, target_indices, product_state):
    """adapters.py:251-277."""
    n = len(state_dims)
    perm, _, reordered = _reorder_meta(tuple(state_dims), tuple(target_indices))
    if perm == list(range(n)):
        return product_state, list(state_dims), perm
    ps = np.asarray(product_state).reshape((*state_dims, *state_dims))
    ps = np.transpose(ps, perm + [p + n for p in perm])
    return ps.reshape((math.prod(reordered),) * 2), reordered, perm


def restore_matrix_order(state_dims, perm, product_state):
    """adapters.py:280-296."""
    if list(perm) == list(range(len(perm))):
        return product_state
    inv = _inverse_perm(perm)
    reordered = _permute_dims(state_dims, perm)
    n = len(inv)
    ps = np.asarray(product_state).reshape((*reordered, *reordered))
    ps = np.transpose(ps, inv + [i + n for i in inv])
    return ps.reshape((math.prod(state_dims),) * 2)


def _check_vec(state_dims, target_indices, product_state, operator):
    """adapters.py:316-327 (ValueError on size, AssertionError on operator shape)."""
    expected = int(math.prod(state_dims))
    if np.size(product_state) != expected:
        raise ValueError(
            f"product_state has size {np.size(product_state)}, expected {expected} "
            f"for dims {state_dims}"
        )
    td = int(math.prod([state_dims[i] for i in target_indices]))
    if tuple(np.shape(operator)) != (td, td):
        raise AssertionError(
            f"Operator shape {np.shape(operator)} does not match expected {(td, td)}"
        )


def _check_mat(state_dims, target_indices, product_state, operator):
    """adapters.py:364-375."""
    expected = int(math.prod(state_dims))
    if np.size(product_state) != expected * expected:
        raise ValueError(
            f"product_state has size {np.size(product_state)}, expected {expected**2} "
            f"for dims {state_dims}"
        )
    td = int(math.prod([state_dims[i] for i in target_indices]))
    if tuple(np.shape(operator)) != (td, td):
        raise AssertionError(
            f"Operator shape {np.shape(operator)} does not match expected {(td, td)}"
        )


def apply_operation_vector(
    state_dims, target_indices, product_state, operator, use_contraction=False
):
    """adapters.py:299-345."""
    target_indices = tuple(target_indices)
    state_dims = tuple(state_dims)
    _check_vec(state_dims, target_indices, product_state, operator)
    if use_contraction:
        td = tuple(state_dims[i] for i in target_indices)
        operands, out = es.apply_operator_vector(len(state_dims), target_indices)
        res = np.einsum(
            np.asarray(operator).reshape((*td, *td)),
            operands[0],
            np.asarray(product_state).reshape((*state_dims, 1)),
            operands[1],
            out,
            optimize=True,
        )
        return res.reshape((-1, 1))
    if target_indices == tuple(range(len(target_indices))):
        return kernels.apply_op_vector_ordered(
            state_dims, len(target_indices), product_state, operator
        )
    ps, reordered, perm = reorder_vector_front(state_dims, target_indices, product_state)
    out = kernels.apply_op_vector_ordered(
        tuple(reordered), len(target_indices), ps, operator
    )
    return restore_vector_order(state_dims, perm, out)


def apply_operation_matrix(
    state_dims, target_indices, product_state, operator, use_contraction=False
):
    """adapters.py:348-393."""
    target_indices = tuple(target_indices)
    state_dims = tuple(state_dims)
    _check_mat(state_dims, target_indices, product_state, operator)
    if use_contraction:
        td = tuple(state_dims[i] for i in target_indices)
        flat = int(math.prod(state_dims))
        operands, out = es.apply_operator_matrix(len(state_dims), target_indices)
        op_t = np.asarray(operator).reshape((*td, *td))
        res = np.einsum(
            op_t,
            operands[0],
            np.asarray(product_state).reshape((*state_dims, *state_dims)),
            operands[1],
            np.conj(op_t),
            operands[2],
            out,
            optimize=True,
        )
        return res.reshape((flat, flat))
    if target_indices == tuple(range(len(target_indices))):
        return kernels.apply_op_matrix_ordered(
            state_dims, len(target_indices), product_state, operator
        )
    ps, reordered, perm = reorder_matrix_front(state_dims, target_indices, product_state)
    out = kernels.apply_op_matrix_ordered(
        tuple(reordered), len(target_indices), ps, operator
    )
    return restore_matrix_order(state_dims, perm, out)


def apply_kraus_matrix(
    state_dims, target_indices, product_state, operators, use_contraction=False
):
    """adapters.py:396-466 (no shape validation in the reference)."""
    target_indices = tuple(target_indices)
    state_dims = tuple(state_dims)
    operators = np.asarray(operators)
    if use_contraction:
        td = [state_dims[i] for i in target_indices]
        flat = int(math.prod(state_dims))
        op_t = operators.reshape((-1, *td, *td))
        operands, out = es.kraus_matrix(len(state_dims), target_indices)
        res = np.einsum(
            op_t,
            operands[0],
            np.asarray(product_state).reshape((*state_dims, *state_dims)),
            operands[1],
            np.conj(op_t),
            operands[2],
            out,
            optimize=True,
        )
        return res.reshape((flat, flat))
    if target_indices == tuple(range(len(target_indices))):
        return kernels.apply_kraus_matrix_ordered(
            state_dims, len(target_indices), product_state, operators
        )
    ps, reordered, perm = reorder_matrix_front(state_dims, target_indices, product_state)
    out = kernels.apply_kraus_matrix_ordered(
        tuple(reordered), len(target_indices), ps, operators
    )
    return restore_matrix_order(state_dims, perm, out)


def _measure(kernel, reorder, is_matrix, state_dims, target_indices, product_state, *key):
    """Shared shape of adapters.py:469-673: reorder, ordered kernel, reshape post."""
    target_indices = tuple(target_indices)
    state_dims = tuple(state_dims)
    _, rest, _ = _reorder_meta(state_dims, target_indices)
    if target_indices == tuple(range(len(target_indices))):
        res = kernel(state_dims, len(target_indices), product_state, *key)
    else:
        ps, reordered, _ = reorder(state_dims, target_indices, product_state)
        res = kernel(tuple(reordered), len(target_indices), ps, *key)
    rest_flat = math.prod(_permute_dims(state_dims, rest))
    shape = (rest_flat, rest_flat) if is_matrix else (rest_flat, 1)
    return res, shape


def measure_vector(state_dims, target_indices, product_state, key):
    """adapters.py:469-501 -> (outcome, post (r,1), key)."""
    (o, post, key), shape = _measure(
        kernels.measure_vector_ordered, reorder_vector_front, False,
        state_dims, target_indices, product_state, key,
    )
    return o, post.reshape(shape), key


def measure_matrix(state_dims, target_indices, product_state, key):
    """adapters.py:504-535 -> (outcome, post (r,r), key)."""
    (o, post, key), shape = _measure(
        kernels.measure_matrix_ordered, reorder_matrix_front, True,
        state_dims, target_indices, product_state, key,
    )
    return o, post.reshape(shape), key


def measure_vector_with_probs(state_dims, target_indices, product_state, key):
    """adapters.py:538-569."""
    (o, post, probs, key), shape = _measure(
        kernels.measure_vector_ordered_with_probs, reorder_vector_front, False,
        state_dims, target_indices, product_state, key,
    )
    return o, post.reshape(shape), probs, key


def measure_matrix_with_probs(state_dims, target_indices, product_state, key):
    """adapters.py:572-613."""
    (o, post, probs, key), shape = _measure(
        kernels.measure_matrix_ordered_with_probs, reorder_matrix_front, True,
        state_dims, target_indices, product_state, key,
    )
    return o, post.reshape(shape), probs, key


def measure_vector_expectation(state_dims, target_indices, product_state):
    """adapters.py:616-643 -> (probs, expected post rho of the rest (r,r))."""
    (probs, post), shape = _measure(
        kernels.measure_vector_expectation_ordered, reorder_vector_front, True,
        state_dims, target_indices, product_state,
    )
    return probs, post.reshape(shape)


def measure_matrix_expectation(state_dims, target_indices, product_state):
    """adapters.py:646-673."""
    (probs, post), shape = _measure(
        kernels.measure_matrix_expectation_ordered, reorder_matrix_front, True,
        state_dims, target_indices, product_state,
    )
    return probs, post.reshape(shape)


def measure_povm_matrix(state_dims, target_indices, operators, product_state, key):
    """adapters.py:676-693: full-size post state, original order restored."""
    ps, reordered, perm = reorder_matrix_front(
        tuple(state_dims), tuple(target_indices), product_state
    )
    outcome, post, key = kernels.measure_povm_matrix_ordered(
        tuple(reordered), len(target_indices), operators, ps, key
    )
    return outcome, restore_matrix_order(state_dims, perm, post), key


def trace_out_matrix(state_dims, target_indices, product_state, use_contraction=False):
    """
    adapters.py:696-733.  ``target_indices`` are the subsystems to KEEP.  The
    ordered route keeps them in the GIVEN order (perm = rest + targets, :720);
    the contraction route keeps them in TENSOR order (:80-81).
    """
    n = len(state_dims)
    rest = [i for i in range(n) if i not in target_indices]
    if len(rest) == 0:
        return product_state
    expected = int(math.prod(state_dims))
    if np.size(product_state) != expected * expected:
        raise ValueError(
            f"product_state has size {np.size(product_state)}, expected {expected**2} "
            f"for dims {state_dims}"
        )
    if use_contraction:
        keep = [i for i in range(n) if i in target_indices]
        operands, out = es.trace_out_matrix(n, keep)
        flat = math.prod(_permute_dims(state_dims, target_indices))
        res = np.einsum(
            np.asarray(product_state).reshape((*state_dims, *state_dims)),
            operands[0],
            out,
        )
        return res.reshape((flat, flat))
    perm = rest + list(target_indices)
    reordered = _permute_dims(state_dims, perm)
    ps = np.asarray(product_state).reshape((*state_dims, *state_dims))
    ps = np.transpose(ps, perm + [p + n for p in perm])
    ps = ps.reshape((math.prod(reordered),) * 2)
    reduced = kernels.trace_out_matrix_ordered(tuple(reordered), len(rest), ps)
    flat = math.prod(_permute_dims(state_dims, target_indices))
    return reduced.reshape((flat, flat))


def trace_out_vector(state_dims, target_indices, product_state, use_contraction=False):
    """state/utils/trace_out.py:17-25: expand psi -> psi psi^dagger, then trace_out_matrix."""
    psi = np.asarray(product_state).reshape((-1, 1))
    rho = psi @ np.conj(psi.T)
    return trace_out_matrix(state_dims, target_indices, rho, use_contraction)
