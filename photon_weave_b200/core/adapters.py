"""
Drop-in for ``photon_weave.core.adapters`` (reference: core/adapters.py): the same
function names, argument order, return shapes and error behaviour, computed by
the sm_100a kernels in libpwb200 instead of opt_einsum/XLA.

Differences a maintainer should know:
* ``use_contraction`` selects between two *numerically equivalent* reference
  routes (einsum on the unpermuted tensor vs transpose -> ordered kernel ->
  transpose back).  Here both run the same strided kernel -- the state is never
  transposed -- so the flag only changes ``trace_out_matrix``'s documented
  ordering of the kept subsystems (given order vs tensor order,
  core/adapters.py:720,732 vs :80-81).
* Arrays are CUDA ``torch.Tensor`` / DLPack device arrays (zero copy) or host
  NumPy arrays (copied in and out).  Outcomes stay on the device (0-d int64
  tensor) unless the inputs were host arrays.
* The ``*_jit`` / ``*_jit_meta`` variants exist for API compatibility; there is
  nothing to trace -- launches are already asynchronous on the current stream.
* Transform rules (core/transforms.py): when an input tensor requires a gradient the call is
  recorded for reverse-mode differentiation (the derivative of an apply is another apply);
  under ``transforms.vmap`` a leading batch axis of the state is served by one launch.
"""

from __future__ import annotations

import math
from typing import Any, List, Sequence, Tuple

import torch

from . import _arrays as A
from . import ffi
from . import transforms as T
from .meta import DimsMeta


from functools import lru_cache


@lru_cache(maxsize=4096)
def _geom(state_dims: Tuple[int, ...], target_indices: Tuple[int, ...]) -> Tuple[int, int]:
    """(prod(dims), prod(target dims)) -- cached: the small configs are host-latency bound."""
    return (int(math.prod(state_dims)), int(math.prod([state_dims[i] for i in target_indices])))


def _rest(state_dims: Sequence[int], target_indices: Sequence[int]) -> List[int]:
    return [i for i in range(len(state_dims)) if i not in target_indices]


def _shape_of(x: Any) -> Tuple[int, ...]:
    return tuple(x.shape) if hasattr(x, "shape") else tuple(torch.as_tensor(x).shape)


def _size_of(x: Any) -> int:
    if type(x) is torch.Tensor:
        return x.numel()
    return int(math.prod(_shape_of(x)))


def _prep(state: Any, *others: Any):
    """Move everything to the device in one common complex dtype."""
    if type(state) is torch.Tensor and state.is_cuda and state.is_complex() and state.is_contiguous() and \
            all(type(o) is torch.Tensor and o.is_cuda and o.dtype == state.dtype and o.is_contiguous() for o in others):
        return False, state, list(others)   # the common case: nothing to move or promote
    A.require_cuda()
    host = A.is_host(state)
    st = A.to_device(state)
    ot = [A.to_device(o) for o in others]
    dt = A.complex_dtype_for(st, *ot)
    return host, A.as_complex(st, dt), [A.as_complex(o, dt) for o in ot]


# --------------------------------------------------------------------------------
# apply (core/adapters.py:299-466)
# --------------------------------------------------------------------------------


def apply_operation_vector(
    state_dims: Sequence[int],
    target_indices: Sequence[int],
    product_state: Any,
    operator: Any,
    use_contraction: bool = False,
):
    """core/adapters.py:299-345 -> (N, 1)."""
    target_indices = tuple(target_indices)
    state_dims = tuple(state_dims)
    expected, target_dim = _geom(state_dims, target_indices)
    batch = T.current_batch()
    if _size_of(product_state) != expected * (batch or 1):
        raise ValueError(
            f"product_state has size {_size_of(product_state)}, expected {expected} "
            f"for dims {state_dims}"
        )
    if _shape_of(operator) != (target_dim, target_dim):
        raise AssertionError(
            f"Operator shape {_shape_of(operator)} does not match expected "
            f"{(target_dim, target_dim)}"
        )
    host, ps, (op,) = _prep(product_state, operator)
    dims, tgt, shape = state_dims, target_indices, (expected, 1)
    if batch:  # the batch axis is one more untouched subsystem: one launch for the whole batch
        dims, tgt, shape = (batch,) + state_dims, tuple(t + 1 for t in target_indices), (batch, expected, 1)
    if T.needs_grad(ps, op):
        out = T.ApplyVec.apply(ps, op, dims, tgt)
    else:
        out = ffi.apply_vec(ps, op, dims, tgt)
    if out.shape != shape:   # (N, 1) in, (N, 1) out is the common case: no view to build
        out = out.reshape(shape)
    return A.back(out, host)


def _two_sided(ps, ops, state_dims, target_indices, single: bool):
    """sum_k K rho K^H on the device: plain, differentiable, or batched (leading batch axis)."""
    expected = int(math.prod(state_dims))
    n = len(state_dims)
    batch = T.current_batch()
    if batch:
        if T.needs_grad(ps, ops):
            raise T.NotBatchable("differentiating a batched two-sided apply")
        axes = (batch,) + state_dims + state_dims
        rows = [1 + t for t in target_indices]
        cols = [1 + n + t for t in target_indices]
        out = ffi.kraus_axes(ps.reshape(-1), ops, axes, rows, cols)
        return out.reshape(batch, expected, expected)
    if T.needs_grad(ps, ops):
        return T.KrausMat.apply(ps, ops, state_dims, target_indices).reshape(expected, expected)
    if single:
        return ffi.apply_mat(ps, ops.reshape(ops.shape[-2:]), state_dims, target_indices).reshape(expected, expected)
    return ffi.kraus_mat(ps, ops, state_dims, target_indices).reshape(expected, expected)


def apply_operation_matrix(
    state_dims: Sequence[int],
    target_indices: Sequence[int],
    product_state: Any,
    operator: Any,
    use_contraction: bool = False,
):
    """core/adapters.py:348-393 -> (N, N)."""
    target_indices = tuple(target_indices)
    state_dims = tuple(state_dims)
    expected = int(math.prod(state_dims))
    if _size_of(product_state) != expected * expected * (T.current_batch() or 1):
        raise ValueError(
            f"product_state has size {_size_of(product_state)}, expected {expected**2} "
            f"for dims {state_dims}"
        )
    target_dim = int(math.prod([state_dims[i] for i in target_indices]))
    if _shape_of(operator) != (target_dim, target_dim):
        raise AssertionError(
            f"Operator shape {_shape_of(operator)} does not match expected "
            f"{(target_dim, target_dim)}"
        )
    host, ps, (op,) = _prep(product_state, operator)
    out = _two_sided(ps, op.reshape(1, target_dim, target_dim), state_dims, target_indices, single=True)
    return A.back(out, host)


def apply_kraus_matrix(
    state_dims: Sequence[int],
    target_indices: Sequence[int],
    product_state: Any,
    operators: Any,
    use_contraction: bool = False,
):
    """core/adapters.py:396-466 -> (N, N); operators stacked (K, t, t)."""
    target_indices = tuple(target_indices)
    state_dims = tuple(state_dims)
    expected = int(math.prod(state_dims))
    target_dim = int(math.prod([state_dims[i] for i in target_indices]))
    if isinstance(operators, (list, tuple)):
        operators = torch.stack([A.to_device(o) for o in operators])
    host, ps, (ops,) = _prep(product_state, operators)
    # the reference reshapes (-1, *target_dims, *target_dims) (core/adapters.py:443)
    ops = ops.reshape(-1, target_dim, target_dim)
    if ps.numel() != expected * expected * (T.current_batch() or 1):
        raise ValueError(
            f"product_state has size {ps.numel()}, expected {expected**2} for dims {state_dims}"
        )
    return A.back(_two_sided(ps, ops, state_dims, target_indices, single=False), host)


# --------------------------------------------------------------------------------
# measurement (core/adapters.py:469-693)
# --------------------------------------------------------------------------------


def _measure(state_dims, target_indices, product_state, key, is_matrix: bool, rescale: bool = True):
    if T.current_batch():
        raise T.NotBatchable("sampling has no batching rule: one call per batch element")
    target_indices = tuple(target_indices)
    state_dims = tuple(state_dims)
    host, ps, _ = _prep(product_state)
    probs = ffi.probs(ps, state_dims, target_indices, is_matrix)
    outcome = ffi.draw(probs, key, ps.dtype)
    post = ffi.collapse(ps, outcome, probs if rescale else None, state_dims, target_indices, is_matrix)
    if host:
        return int(outcome.item()), post.cpu().numpy(), probs.cpu().numpy()
    return outcome, post, probs


def measure_vector(state_dims, target_indices, product_state, key):
    """core/adapters.py:469-501 -> (outcome, post (r,1), key); key returned unchanged."""
    o, post, _ = _measure(state_dims, target_indices, product_state, key, False)
    return o, post, key


def measure_matrix(state_dims, target_indices, product_state, key):
    """core/adapters.py:504-535 -> (outcome, post (r,r), key)."""
    o, post, _ = _measure(state_dims, target_indices, product_state, key, True)
    return o, post, key


def measure_vector_with_probs(state_dims, target_indices, product_state, key):
    """core/adapters.py:538-569."""
    o, post, probs = _measure(state_dims, target_indices, product_state, key, False)
    return o, post, probs, key


def measure_matrix_with_probs(state_dims, target_indices, product_state, key):
    """core/adapters.py:572-613."""
    o, post, probs = _measure(state_dims, target_indices, product_state, key, True)
    return o, post, probs, key


def measure_vector_expectation(state_dims, target_indices, product_state):
    """
    core/adapters.py:616-643 -> (probs, sum_t p_t |phi_t><phi_t|), the reduced
    density matrix of the unmeasured subsystems (core/kernels.py:266-283),
    computed in one pass without forming the per-outcome branches.
    """
    target_indices = tuple(target_indices)
    state_dims = tuple(state_dims)
    if T.current_batch():
        raise T.NotBatchable("expectation of a batch: one call per batch element")
    host, ps, _ = _prep(product_state)
    rest = _rest(state_dims, target_indices)
    if T.needs_grad(ps):  # differentiable (tests/state/utils/test_measurements.py:337-353)
        return T.ProbsVec.apply(ps, state_dims, target_indices), T.ReducedFromVec.apply(ps, state_dims, tuple(rest))
    probs = ffi.probs(ps, state_dims, target_indices, False)
    post = ffi.reduced_from_vec(ps, state_dims, rest)
    return A.back(probs, host), A.back(post, host)


def measure_matrix_expectation(state_dims, target_indices, product_state):
    """core/adapters.py:646-673 -> (probs, Tr_targets rho)."""
    target_indices = tuple(target_indices)
    state_dims = tuple(state_dims)
    if T.current_batch():
        raise T.NotBatchable("expectation of a batch: one call per batch element")
    host, ps, _ = _prep(product_state)
    rest = _rest(state_dims, target_indices)
    if T.needs_grad(ps):  # differentiable (tests/state/utils/test_measurements.py:356-371)
        return T.ProbsMat.apply(ps, state_dims, target_indices), T.TraceOutMat.apply(ps, state_dims, tuple(rest))
    probs = ffi.probs(ps, state_dims, target_indices, True)
    post = ffi.trace_out_mat(ps, state_dims, rest)
    return A.back(probs, host), A.back(post, host)


def measure_povm_matrix(state_dims, target_indices, operators, product_state, key):
    """
    core/adapters.py:676-693 -> (outcome, post (N,N) in the original order, key).
    Weights come from the reduced state of the targets (p_k = Tr(E_k rho_T E_k^dagger))
    instead of K full-size projections; only the sampled branch is applied.
    """
    target_indices = tuple(target_indices)
    state_dims = tuple(state_dims)
    expected = int(math.prod(state_dims))
    target_dim = int(math.prod([state_dims[i] for i in target_indices]))
    if isinstance(operators, (list, tuple)):
        operators = torch.stack([A.to_device(o) for o in operators])
    host, ps, (ops,) = _prep(product_state, operators)
    ops = ops.reshape(-1, target_dim, target_dim)
    rho_t = ffi.trace_out_mat(ps, state_dims, target_indices)
    probs, traces = ffi.povm_probs(rho_t.contiguous(), ops)
    outcome = ffi.draw(probs, key, ps.dtype)
    op_sel = ffi.povm_select(ops, outcome, traces)
    post = ffi.apply_mat(ps, op_sel, state_dims, target_indices).reshape(expected, expected)
    if host:
        return int(outcome.item()), post.cpu().numpy(), key
    return outcome, post, key


# --------------------------------------------------------------------------------
# partial trace (core/adapters.py:696-733)
# --------------------------------------------------------------------------------


def trace_out_matrix(state_dims, target_indices, product_state, use_contraction: bool = False):
    """
    core/adapters.py:696-733: keep ``target_indices``.  Ordered route: kept
    subsystems in the GIVEN order; contraction route: in TENSOR order.
    """
    state_dims = tuple(state_dims)
    target_indices = tuple(target_indices)
    rest = _rest(state_dims, target_indices)
    if len(rest) == 0:
        return product_state
    expected = int(math.prod(state_dims))
    if _size_of(product_state) != expected * expected:
        raise ValueError(
            f"product_state has size {_size_of(product_state)}, expected {expected**2} "
            f"for dims {state_dims}"
        )
    keep = sorted(target_indices) if use_contraction else list(target_indices)
    if T.current_batch():
        raise T.NotBatchable("partial trace of a batch: one call per batch element")
    host, ps, _ = _prep(product_state)
    if T.needs_grad(ps):
        return T.TraceOutMat.apply(ps, state_dims, tuple(keep))
    return A.back(ffi.trace_out_mat(ps, state_dims, keep), host)


def trace_out_vector(state_dims, target_indices, product_state, use_contraction: bool = False):
    """
    state/utils/trace_out.py:17-25 computes psi psi^dagger (N x N) and traces it;
    here the reduced matrix is accumulated directly from psi in one pass.
    """
    state_dims = tuple(state_dims)
    target_indices = tuple(target_indices)
    keep = tuple(sorted(target_indices)) if use_contraction else target_indices
    if T.current_batch():
        raise T.NotBatchable("partial trace of a batch: one call per batch element")
    host, ps, _ = _prep(product_state)
    if T.needs_grad(ps):
        return T.ReducedFromVec.apply(ps, state_dims, keep)
    return A.back(ffi.reduced_from_vec(ps, state_dims, keep), host)


# --------------------------------------------------------------------------------
# *_jit / *_jit_meta surface (core/adapters.py:736-1011)
# --------------------------------------------------------------------------------


def apply_operation_vector_jit(state_dims, target_indices, product_state, operator, use_contraction=False):
    return apply_operation_vector(tuple(state_dims), tuple(target_indices), product_state, operator, use_contraction)


def apply_operation_vector_jit_meta(meta: DimsMeta, product_state, operator, use_contraction=False):
    return apply_operation_vector(meta.state_dims, meta.target_indices, product_state, operator, use_contraction)


def apply_operation_matrix_jit(state_dims, target_indices, product_state, operator, use_contraction=False):
    return apply_operation_matrix(tuple(state_dims), tuple(target_indices), product_state, operator, use_contraction)


def apply_operation_matrix_jit_meta(meta: DimsMeta, product_state, operator, use_contraction=False):
    return apply_operation_matrix(meta.state_dims, meta.target_indices, product_state, operator, use_contraction)


def apply_kraus_matrix_jit(state_dims, target_indices, product_state, operators, use_contraction=False):
    return apply_kraus_matrix(tuple(state_dims), tuple(target_indices), product_state, operators, use_contraction)


def apply_kraus_matrix_jit_meta(meta: DimsMeta, product_state, operators, use_contraction=False):
    return apply_kraus_matrix(meta.state_dims, meta.target_indices, product_state, operators, use_contraction)


def measure_vector_jit(state_dims, target_indices, product_state, key):
    return measure_vector(tuple(state_dims), tuple(target_indices), product_state, key)


def measure_vector_jit_meta(meta: DimsMeta, product_state, key):
    return measure_vector(meta.state_dims, meta.target_indices, product_state, key)


def measure_matrix_jit(state_dims, target_indices, product_state, key):
    return measure_matrix(tuple(state_dims), tuple(target_indices), product_state, key)


def measure_matrix_jit_meta(meta: DimsMeta, product_state, key):
    return measure_matrix(meta.state_dims, meta.target_indices, product_state, key)


def measure_povm_matrix_jit(state_dims, target_indices, operators, product_state, key):
    return measure_povm_matrix(tuple(state_dims), tuple(target_indices), operators, product_state, key)


def measure_povm_matrix_jit_meta(meta: DimsMeta, operators, product_state, key):
    return measure_povm_matrix(meta.state_dims, meta.target_indices, operators, product_state, key)


def trace_out_matrix_jit(state_dims, target_indices, product_state, use_contraction=False):
    return trace_out_matrix(tuple(state_dims), tuple(target_indices), product_state, bool(use_contraction))


def trace_out_matrix_jit_meta(meta: DimsMeta, product_state, use_contraction=False):
    return trace_out_matrix(meta.state_dims, meta.target_indices, product_state, use_contraction=use_contraction)
