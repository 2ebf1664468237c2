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


def ensure_static_dims_for_ops(state_objs: Sequence, operations: Sequence) -> None:
    """Pre-size Fock states for a list of operations before entering static (use_jit)
    mode (shape_planning.py:51-69)."""
    from ...operation.fock_operation import FockOperationType

    for op in operations:
        for so in state_objs:
            if isinstance(op._operation_type, FockOperationType):
                if getattr(so, "_num_quanta", None) is None:
                    continue
                op.compute_dimensions(so._num_quanta, so.trace_out())
                if so.dimensions != op.dimensions[0]:
                    so.resize(op.dimensions[0])


def envelope_plan(envelope, *target_states) -> ShapePlan:
    """Plan over (fock, polarization); no targets = both (shape_planning.py:72-79)."""
    state_objs = [envelope.fock, envelope.polarization]
    return build_plan(state_objs, target_states if target_states else state_objs)


def composite_plan(composite_envelope, product_state_index: int = 0, *target_states) -> ShapePlan:
    """Plan over one product state of a composite envelope (shape_planning.py:82-89)."""
    state_objs = list(composite_envelope.product_states[product_state_index].state_objs)
    return build_plan(state_objs, target_states if target_states else state_objs)


def composite_meta(composite_envelope, product_state_index: int = 0, *target_states) -> DimsMeta:
    return composite_plan(composite_envelope, product_state_index, *target_states).meta


def compiled_kernels(plan):
    """Callables bound to a plan (shape_planning.py:92-133)."""

    if isinstance(plan, DimsMeta):
        plan = ShapePlan(dims=plan.state_dims, target_indices=plan.target_indices,
                         rest_indices=plan.rest_indices, meta=plan)

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
