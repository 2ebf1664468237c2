"""
Partial-trace bridges (reference: photon_weave/state/utils/trace_out.py:17-72).
"""

from __future__ import annotations

from ...core import adapters
from ...photon_weave import Config


def trace_out_vector(state_objs, target_states, product_state):
    """
    trace_out.py:17-25 expands psi to psi psi^dagger (N x N) and traces it; here the reduced
    matrix is accumulated directly from psi (pw_reduced_from_vec) -- same result, O(N) memory.
    """
    objs = list(state_objs)
    dims = [s.dimensions for s in objs]
    target_indices = [objs.index(s) for s in target_states]
    return adapters.trace_out_vector(dims, target_indices, product_state,
                                     use_contraction=bool(Config().contractions))


def trace_out_matrix(state_objs, target_states, product_state, meta=None, use_contraction=None):
    """trace_out.py:28-51."""
    if use_contraction is None:
        use_contraction = Config().contractions
    if meta is None:
        objs = list(state_objs)
        dims = [s.dimensions for s in objs]
        target_indices = [objs.index(s) for s in target_states]
        return adapters.trace_out_matrix(dims, target_indices, product_state,
                                         use_contraction=bool(use_contraction))
    return adapters.trace_out_matrix_jit_meta(meta, product_state, use_contraction=bool(use_contraction))


def trace_out_matrix_jit(state_objs, target_states, product_state, use_contraction=None):
    """trace_out.py:54-72."""
    return trace_out_matrix(state_objs, target_states, product_state, None, use_contraction)
