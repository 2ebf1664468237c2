"""
Objects -> dims bridge for apply / Kraus (reference: photon_weave/state/utils/operations.py).
"""

from __future__ import annotations

from math import prod
from typing import Optional, Sequence, Tuple

from ...core import adapters
from ...core.meta import DimsMeta
from ...photon_weave import Config
from .shape_planning import ShapePlan


def _extract_meta(meta, state_objs, target_states):
    """operations.py:27-45."""
    if isinstance(meta, ShapePlan):
        return meta.dims, meta.target_indices
    if meta is None:
        objs = list(state_objs)
        return tuple(s.dimensions for s in objs), tuple(objs.index(s) for s in target_states)
    return meta.state_dims, meta.target_indices


def _shape(x) -> Tuple[int, ...]:
    return tuple(x.shape)


def apply_operation_vector(state_objs, target_states, product_state, operator, meta=None, key=None,
                           use_contraction: Optional[bool] = None, target_indices=None):
    """operations.py:91-173."""
    use_contraction = Config().contractions if use_contraction is None else use_contraction
    dims, tgt = _extract_meta(meta, state_objs, target_states)
    if target_indices is None:
        target_indices = tgt
    target_dim = int(prod(dims[i] for i in target_indices))
    if _shape(operator) != (target_dim, target_dim):
        raise AssertionError(
            f"Operator shape {_shape(operator)} does not match expected {(target_dim, target_dim)}")
    return adapters.apply_operation_vector(dims, target_indices, product_state, operator,
                                           use_contraction=bool(use_contraction))


def apply_operation_matrix(state_objs, target_states, product_state, operator, meta=None, key=None,
                           use_contraction: Optional[bool] = None, target_indices=None):
    """operations.py:176-257."""
    use_contraction = Config().contractions if use_contraction is None else use_contraction
    dims, tgt = _extract_meta(meta, state_objs, target_states)
    if target_indices is None:
        target_indices = tgt
    target_dim = int(prod(dims[i] for i in target_indices))
    if _shape(operator) != (target_dim, target_dim):
        raise AssertionError(
            f"Operator shape {_shape(operator)} does not match expected {(target_dim, target_dim)}")
    return adapters.apply_operation_matrix(dims, target_indices, product_state, operator,
                                           use_contraction=bool(use_contraction))


def apply_kraus_matrix(state_objs, target_states, product_state, operators: Sequence, meta=None,
                       use_contraction: Optional[bool] = None, target_indices=None):
    """operations.py:260-337: validates each operator, stacks, applies."""
    if use_contraction is None:
        use_contraction = Config().contractions
    dims, tgt = _extract_meta(meta, state_objs, target_states)
    if target_indices is None:
        target_indices = tgt
    target_dim = int(prod(dims[i] for i in target_indices))
    expected = (target_dim, target_dim)
    for op in operators:
        if _shape(op) != expected:
            raise AssertionError(f"Kraus operator shape {_shape(op)} does not match expected {expected}")
    return adapters.apply_kraus_matrix(dims, target_indices, product_state, list(operators),
                                       use_contraction=bool(use_contraction))
