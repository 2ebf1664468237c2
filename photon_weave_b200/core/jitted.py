"""
DimsMeta entry points (reference: photon_weave/core/jitted.py:19-110).
"""

from __future__ import annotations

from . import adapters
from .meta import DimsMeta


def apply_operation_vector(meta: DimsMeta, product_state, operator, use_contraction=False):
    return adapters.apply_operation_vector_jit_meta(meta, product_state, operator, use_contraction)


def apply_operation_matrix(meta: DimsMeta, product_state, operator, use_contraction=False):
    return adapters.apply_operation_matrix_jit_meta(meta, product_state, operator, use_contraction)


def apply_kraus_matrix(meta: DimsMeta, product_state, operators, use_contraction=False):
    return adapters.apply_kraus_matrix_jit_meta(meta, product_state, operators, use_contraction=use_contraction)


def measure_vector(meta: DimsMeta, product_state, key):
    return adapters.measure_vector_jit_meta(meta, product_state, key)


def measure_matrix(meta: DimsMeta, product_state, key):
    return adapters.measure_matrix_jit_meta(meta, product_state, key)


def measure_povm_matrix(meta: DimsMeta, operators, product_state, key):
    return adapters.measure_povm_matrix_jit_meta(meta, operators, product_state, key)


def trace_out_matrix(meta: DimsMeta, product_state, use_contraction=False):
    return adapters.trace_out_matrix_jit_meta(meta, product_state, use_contraction=use_contraction)


def measure_vector_with_probs(meta: DimsMeta, product_state, key):
    return adapters.measure_vector_with_probs(meta.state_dims, meta.target_indices, product_state, key)


def measure_matrix_with_probs(meta: DimsMeta, product_state, key):
    return adapters.measure_matrix_with_probs(meta.state_dims, meta.target_indices, product_state, key)


def measure_vector_expectation(meta: DimsMeta, product_state):
    return adapters.measure_vector_expectation(meta.state_dims, meta.target_indices, product_state)


def measure_matrix_expectation(meta: DimsMeta, product_state):
    return adapters.measure_matrix_expectation(meta.state_dims, meta.target_indices, product_state)
