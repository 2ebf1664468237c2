"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU baseline for bench.py (``cpu_baseline`` and ``--impl reference`` legs): the
reference's apply path restated the way XLA-CPU executes it -- one batched
``dot_general`` per operation, no materialised transposes -- using NumPy's
BLAS-backed ``matmul`` so all host cores take part (kind = "port"; the real
reference needs jax/jaxlib, which are not installed in this image).

Follows core/kernels.py:62-98 ("ab,bcn->acn") on the (L, d, R) view the einsum
string of core/adapters.py:112-125 contracts over.
"""

from __future__ import annotations

import math
from typing import Sequence

import numpy as np


def apply_vector_inplace_view(dims: Sequence[int], targets: Sequence[int], psi: np.ndarray,
                              op: np.ndarray) -> np.ndarray:
    """psi' = (op on targets) psi for one target axis or two ADJACENT ascending axes."""
    dims = tuple(dims)
    targets = tuple(targets)
    if len(targets) == 2 and targets[1] == targets[0] + 1:
        i = targets[0]
        d = dims[i] * dims[i + 1]
        L = math.prod(dims[:i])
        R = math.prod(dims[i + 2:])
    elif len(targets) == 1:
        i = targets[0]
        d = dims[i]
        L = math.prod(dims[:i])
        R = math.prod(dims[i + 1:])
    else:
        from . import adapters_np

        return adapters_np.apply_operation_vector(dims, targets, psi, op, True)
    x = psi.reshape(L, d, R)
    if R == 1:
        return (x.reshape(L, d) @ op.T).reshape(-1, 1)
    return np.matmul(op, x).reshape(-1, 1)


def run_layer(dims, layer, psi):
    """Apply a list of (targets, operator) to psi (N,1); returns the new state."""
    for targets, op in layer:
        psi = apply_vector_inplace_view(dims, targets, psi, op)
    return psi
