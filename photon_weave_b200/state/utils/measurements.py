"""
Projective / POVM measurement bridges (reference: photon_weave/state/utils/measurements.py:20-682).
The PNR detector-noise helpers (measurements.py:685-1112) are out of scope.

Key threading follows the reference exactly:
* legacy sequential paths: one key per target from ``_key_sequence`` (``split(key, n+1)[1:]``),
* jit paths: ``use_key, next_key = borrow_key(key)`` then ONE joint draw over prod(target dims)
  and a mixed-radix decode of the flat outcome (measurements.py:200-208).
"""

from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from ...core import adapters
from ...core import rng as _rng
from ...core.meta import DimsMeta
from ...core.rng import borrow_key
from ...photon_weave import Config
from .shape_planning import ShapePlan


def _ensure_key(key):
    """measurements.py:20-24."""
    return key if key is not None else Config().random_key


def _borrow_optional(key):
    """measurements.py:27-47."""
    if key is None:
        return Config().random_key, None
    k = _rng.split(key, 2)
    return k[0], k[1]


def _key_sequence(key, count: int):
    """measurements.py:50-80: drop the first subkey, return (keys, leftover)."""
    if key is None:
        cfg = Config()
        return [cfg.random_key for _ in range(count)], None
    k = _rng.split(key, count + 1)
    return [k[i] for i in range(1, count + 1)], k[0]


def _dims_targets(meta, state_objs, target_states, target_indices):
    if isinstance(meta, ShapePlan):
        dims, tgt = list(meta.dims), list(meta.target_indices)
    elif meta is None:
        objs = list(state_objs)
        dims = [s.dimensions for s in objs]
        tgt = [i for i, s in enumerate(objs) if s in target_states]
    else:
        dims, tgt = list(meta.state_dims), list(meta.target_indices)
    if target_indices is not None:
        tgt = list(target_indices)
    return dims, tgt


def _decode(outcome, dims, tgt, target_states) -> Dict:
    """Mixed-radix decode of the joint outcome (measurements.py:200-208)."""
    tmp = int(outcome)
    decoded = []
    for dim in reversed([dims[i] for i in tgt]):
        decoded.append(tmp % dim)
        tmp //= dim
    decoded.reverse()
    return {ts: int(v) for ts, v in zip(target_states, decoded)}


def _sequential(state_objs, target_states, product_state, prob_callback, key, is_matrix: bool):
    """Legacy path (measurements.py:83-138, 212-281): targets measured one at a time."""
    objs = list(state_objs)
    keys, _ = _key_sequence(key, len(target_states))
    outcomes: Dict = {}
    for state, use_key in zip(target_states, keys):
        dims = [s.dimensions for s in objs]
        idx = objs.index(state)
        # jnp.take / plain slicing in the reference: the collapsed slice is NOT rescaled
        outcome, product_state, probs = adapters._measure(dims, (idx,), product_state, use_key,
                                                           is_matrix, rescale=False)
        if prob_callback is not None:
            prob_callback(state, probs)
        outcomes[state] = int(outcome)
        objs.remove(state)
    return outcomes, product_state


def measure_vector(state_objs, target_states, product_state, prob_callback: Optional[Callable] = None,
                   key=None, target_indices=None):
    """measurements.py:83-138 (the collapsed slice is returned un-normalised, as jnp.take does)."""
    return _sequential(state_objs, target_states, product_state, prob_callback, key, False)


def measure_matrix(state_objs, target_states, product_state, prob_callback: Optional[Callable] = None,
                   key=None, target_indices=None):
    """measurements.py:212-281."""
    return _sequential(state_objs, target_states, product_state, prob_callback, key, True)


def measure_vector_jit(state_objs, target_states, product_state, key, meta=None, target_indices=None):
    """measurements.py:141-209 -> (outcomes, post, next_key)."""
    use_key, next_key = borrow_key(key)
    dims, tgt = _dims_targets(meta, state_objs, target_states, target_indices)
    outcome, post, _ = adapters.measure_vector(dims, tgt, product_state, use_key)
    return _decode(outcome, dims, tgt, target_states), post, next_key


def measure_matrix_jit(state_objs, target_states, product_state, key, meta=None, target_indices=None):
    """measurements.py:284-343."""
    use_key, next_key = borrow_key(key)
    dims, tgt = _dims_targets(meta, state_objs, target_states, target_indices)
    outcome, post, _ = adapters.measure_matrix(dims, tgt, product_state, use_key)
    return _decode(outcome, dims, tgt, target_states), post, next_key


def measure_vector_jit_with_probs(state_objs, target_states, product_state, key, meta=None,
                                  target_indices=None):
    """measurements.py:346-417."""
    use_key, next_key = borrow_key(key)
    dims, tgt = _dims_targets(meta, state_objs, target_states, target_indices)
    outcome, post, probs, _ = adapters.measure_vector_with_probs(dims, tgt, product_state, use_key)
    return _decode(outcome, dims, tgt, target_states), post, probs, next_key


def measure_matrix_jit_with_probs(state_objs, target_states, product_state, key, meta=None,
                                  target_indices=None):
    """measurements.py:420-491."""
    use_key, next_key = borrow_key(key)
    dims, tgt = _dims_targets(meta, state_objs, target_states, target_indices)
    outcome, post, probs, _ = adapters.measure_matrix_with_probs(dims, tgt, product_state, use_key)
    return _decode(outcome, dims, tgt, target_states), post, probs, next_key


def measure_vector_expectation(state_objs, target_states, product_state, meta=None):
    """measurements.py:494-534."""
    dims, tgt = _dims_targets(meta, state_objs, target_states, None)
    return adapters.measure_vector_expectation(dims, tgt, product_state)


def measure_matrix_expectation(state_objs, target_states, product_state, meta=None):
    """measurements.py:537-577."""
    dims, tgt = _dims_targets(meta, state_objs, target_states, None)
    return adapters.measure_matrix_expectation(dims, tgt, product_state)


def measure_POVM_matrix(state_objs, target_states, operators, product_state, key=None, meta=None):
    """measurements.py:580-645 -> (outcome index, post state)."""
    key = _ensure_key(key)
    if meta is None:
        objs = list(state_objs)
        dims = [s.dimensions for s in objs]
        tgt = [objs.index(s) for s in target_states]
    else:
        dims, tgt = list(meta.state_dims), list(meta.target_indices)
    target_dim = int(np.prod([dims[i] for i in tgt]))
    expected = (target_dim, target_dim)
    for op in operators:
        if tuple(op.shape) != expected:
            raise AssertionError(f"POVM operator shape {tuple(op.shape)} does not match expected {expected}")
    use_key, _ = borrow_key(key)
    outcome, post, _ = adapters.measure_povm_matrix(dims, tgt, list(operators), product_state, use_key)
    return outcome, post


def measure_POVM_matrix_jit(state_objs, target_states, operators, product_state, key):
    """measurements.py:648-682 -> (outcome, post, key)."""
    objs = list(state_objs)
    dims = [s.dimensions for s in objs]
    tgt = [objs.index(s) for s in target_states]
    return adapters.measure_povm_matrix(dims, tgt, operators, product_state, key)
