"""Re-exports of the operator helpers plus ``fidelity`` / ``kron_reduce``
(reference: photon_weave/core/ops.py)."""

from __future__ import annotations

from typing import Any, Sequence

import numpy as np

from .._math import ops as _ops
from . import device as dev

identity_operator = _ops.identity_operator
hadamard_operator = _ops.hadamard_operator
x_operator = _ops.x_operator
y_operator = _ops.y_operator
z_operator = _ops.z_operator
s_operator = _ops.s_operator
t_operator = _ops.t_operator
controlled_not_operator = _ops.controlled_not_operator
controlled_z_operator = _ops.controlled_z_operator
swap_operator = _ops.swap_operator
sx_operator = _ops.sx_operator
controlled_swap_operator = _ops.controlled_swap_operator
rx_operator = _ops.rx_operator
ry_operator = _ops.ry_operator
rz_operator = _ops.rz_operator
u3_operator = _ops.u3_operator
displacement_operator = _ops.displacement_operator
phase_operator = _ops.phase_operator
squeezing_operator = _ops.squeezing_operator
creation_operator = _ops.creation_operator
annihilation_operator = _ops.annihilation_operator
number_operator = _ops.number_operator
apply_kraus = _ops.apply_kraus
kraus_identity_check = _ops.kraus_identity_check
num_quanta_vector = _ops.num_quanta_vector
num_quanta_matrix = _ops.num_quanta_matrix
jitted_exp = _ops.jitted_exp


def compute_einsum(einsum_str: str, *operands: Any) -> Any:
    """General (small-array) Einstein summation on the device (core/ops.py:67-87); not on the
    state-evolution path, which never evaluates einsum strings."""
    import torch

    ts = [dev.asarray(o) if not dev.is_array(o) else o for o in operands]
    return torch.einsum(einsum_str, *ts)


def fidelity(a: Any, b: Any) -> float:
    """|<a|b>|^2 for two (small) state vectors (core/ops.py:43-49)."""
    return float(np.abs(np.vdot(dev.to_numpy(a), dev.to_numpy(b))) ** 2)


def kron_reduce(arrays: Sequence[Any]) -> Any:
    """Left fold of Kronecker products on the device (core/ops.py:52-64)."""
    if not arrays:
        return dev.scalar_one()
    out = arrays[0]
    for nxt in arrays[1:]:
        out = dev.kron(out, nxt)
    return out
