"""
Reusable dimension metadata (reference: photon_weave/state/utils/shape_planning.py:22-133).
State objects are duck-typed: anything with ``.dimensions`` (the reference's own tests
use ``DummyState(dimensions)``, tests/state/test_shape_planning.py:14-16).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence, Tuple

from ...core import jitted
from ...core.meta import DimsMeta, make_meta
from ...photon_weave import Config


@dataclass(frozen=True)
class ShapePlan:
    dims: Tuple[int, ...]
    target_indices: Tuple[int, ...]
    rest_indices: Tuple[int, ...]
    meta: DimsMeta


def build_plan(state_objs: Sequence, target_states: Sequence) -> ShapePlan:
    """shape_planning.py:29-40."""
    dims = tuple(s.dimensions for s in state_objs)
    target_indices = tuple(list(state_objs).index(s) for s in target_states)
    meta = make_meta(dims, target_indices)
    return ShapePlan(dims=dims, target_indices=target_indices, rest_indices=meta.rest_indices, meta=meta)


def build_meta(state_objs: Sequence, target_states: Sequence) -> DimsMeta:
    """shape_planning.py:43-48."""
    dims = tuple(s.dimensions for s in state_objs)
    target_indices = tuple(list(state_objs).index(s) for s in target_states)
    return make_meta(dims, target_indices)


def compiled_kernels(plan: ShapePlan):
    """Callables bound to a plan (shape_planning.py:92-133)."""

    class Kernels:
        @staticmethod
        def apply_op_vector(ps, op, use_contraction=False):
            return jitted.apply_operation_vector(plan.meta, ps, op, use_contraction=use_contraction)

        @staticmethod
        def apply_op_matrix(ps, op, use_contraction=False):
            return jitted.apply_operation_matrix(plan.meta, ps, op, use_contraction=use_contraction)

        @staticmethod
        def apply_kraus_matrix(ps, ops):
            return jitted.apply_kraus_matrix(plan.meta, ps, ops, use_contraction=Config().contractions)

        @staticmethod
        def measure_vector(ps, key):
            return jitted.measure_vector(plan.meta, ps, key)

        @staticmethod
        def measure_matrix(ps, key):
            return jitted.measure_matrix(plan.meta, ps, key)

        @staticmethod
        def measure_povm_matrix(ops, ps, key):
            return jitted.measure_povm_matrix(plan.meta, ops, ps, key)

        @staticmethod
        def trace_out_matrix(ps, use_contraction=False):
            return jitted.trace_out_matrix(plan.meta, ps, use_contraction=use_contraction)

    return Kernels()
