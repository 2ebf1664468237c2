"""
Positional-dims pass-throughs (reference: photon_weave/core/linear.py:15-106).
"""

from __future__ import annotations

from . import adapters


def apply_operation_vector(state_dims, target_indices, product_state, operator, key=None, use_contraction=False):
    return adapters.apply_operation_vector(state_dims, target_indices, product_state, operator, use_contraction=use_contraction)


def apply_operation_matrix(state_dims, target_indices, product_state, operator, key=None, use_contraction=False):
    return adapters.apply_operation_matrix(state_dims, target_indices, product_state, operator, use_contraction=use_contraction)


def measure_vector(state_dims, target_indices, product_state, key):
    return adapters.measure_vector(state_dims, target_indices, product_state, key)


def measure_matrix(state_dims, target_indices, product_state, key):
    return adapters.measure_matrix(state_dims, target_indices, product_state, key)


def measure_povm_matrix(state_dims, target_indices, operators, product_state, key):
    return adapters.measure_povm_matrix(state_dims, target_indices, operators, product_state, key)


def trace_out_matrix(state_dims, target_indices, product_state, use_contraction=False):
    return adapters.trace_out_matrix(state_dims, target_indices, product_state, use_contraction=use_contraction)


def apply_kraus_matrix(state_dims, target_indices, product_state, operators, use_contraction=False):
    return adapters.apply_kraus_matrix(state_dims, target_indices, product_state, operators, use_contraction=use_contraction)
