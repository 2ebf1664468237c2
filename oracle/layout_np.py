"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

NumPy restatement of the layout / representation steps either side of an apply
(SURVEY.md section 8f.2).
"""

from __future__ import annotations

import numpy as np


def resize_fock_vector(dims, axis, new_dim, psi):
    """state/composite_envelope.py:513-542 (pad with zeros / slice along one subsystem)."""
    t = np.asarray(psi).reshape(dims)
    if new_dim >= dims[axis]:
        pad = [(0, 0)] * len(dims)
        pad[axis] = (0, new_dim - dims[axis])
        t = np.pad(t, pad, mode="constant")
    else:
        sl = [slice(None)] * len(dims)
        sl[axis] = slice(0, new_dim)
        t = t[tuple(sl)]
    return np.ascontiguousarray(t).reshape(-1, 1)


def resize_fock_matrix(dims, axis, new_dim, rho):
    """The intended effect of state/composite_envelope.py:543-585: pad / slice the subsystem on
    BOTH the row and the column side."""
    n = len(dims)
    t = np.asarray(rho).reshape(tuple(dims) + tuple(dims))
    if new_dim >= dims[axis]:
        pad = [(0, 0)] * (2 * n)
        pad[axis] = pad[n + axis] = (0, new_dim - dims[axis])
        t = np.pad(t, pad, mode="constant")
    else:
        sl = [slice(None)] * (2 * n)
        sl[axis] = sl[n + axis] = slice(0, new_dim)
        t = t[tuple(sl)]
    nd = list(dims)
    nd[axis] = new_dim
    N = int(np.prod(nd))
    return np.ascontiguousarray(t).reshape(N, N)


def state_expand_vector(psi):
    """state/utils/state_transform.py:58-63."""
    v = np.asarray(psi).reshape(-1)
    return np.outer(v, np.conj(v))


def state_contract_matrix(rho, tol=1e-6):
    """state/utils/state_transform.py:122-137 -> (state, success)."""
    rho = np.asarray(rho)
    if abs(np.trace(rho @ rho) - 1) < tol:
        w, v = np.linalg.eigh(rho)
        psi = v[:, int(np.argmax(w))].reshape(-1, 1)
        if abs(psi[0, 0]) > 1e-12:
            psi = psi * np.exp(-1j * np.angle(psi[0]))
        else:
            # psi[0] == 0: the reference applies no phase and keeps whatever eigh returned
            # (backend dependent).  Its own golden (tests/operation/test_composite_operation.py,
            # test_non_polarizing_bs_matrix: +1/sqrt(2) at |0,2> and |2,0>) pins the
            # representative whose largest component (lowest index on ties) is real positive.
            j = int(np.argmax(np.round(np.abs(psi[:, 0]), 12)))
            psi = psi * np.exp(-1j * np.angle(psi[j]))
        return psi, True
    return rho, False


def num_quanta_vector(vector):
    """_math/ops.py:472-486 (highest row index holding a non-zero entry)."""
    return int(np.nonzero(np.asarray(vector))[0][-1])


def num_quanta_matrix(matrix):
    """_math/ops.py:489-512."""
    m = np.asarray(matrix)
    rows = np.where(np.any(m != 0, axis=1))[0]
    cols = np.where(np.any(m != 0, axis=0))[0]
    return int(max(rows[-1], cols[-1]))


def occupancy_levels(dims, state, is_matrix=False):
    """The per-subsystem loop of state/composite_envelope.py:688-696: reduce the product state to
    each subsystem (state/utils/trace_out.py:17-25 for vectors, core/adapters.py:696-733 for
    matrices) and take num_quanta of the reduced density matrix."""
    from . import adapters_np as ad

    out = []
    for i in range(len(dims)):
        if is_matrix:
            red = ad.trace_out_matrix(tuple(dims), (i,), np.asarray(state))
            out.append(num_quanta_matrix(red))
        else:
            red = ad.trace_out_vector(tuple(dims), (i,), np.asarray(state))
            out.append(num_quanta_matrix(red))
    return out
