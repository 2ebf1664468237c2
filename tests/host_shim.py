"""
TEST INFRASTRUCTURE ONLY.

The state containers (Fock / Polarization / CustomState / Envelope / CompositeEnvelope) reach
the GPU through ``photon_weave_b200.core.ffi`` and nothing else.  On a machine without a GPU
the ``-m "not gpu"`` suite still has to cover their HOST logic (routing, ordering, key
threading, bookkeeping), so this module swaps every ``ffi`` entry point the containers use
for the NumPy oracle (``oracle/``) acting on CPU ``torch`` tensors.  The product package never
imports this file; ``tests/test_gpu_containers.py`` runs the very same cases on the real CUDA
path.
"""

from __future__ import annotations

import contextlib
import math

import numpy as np
import torch

from oracle import adapters_np as oad
from oracle import layout_np as olay
from oracle import rng_threefry as orng
from photon_weave_b200.core import _arrays as A
from photon_weave_b200.core import ffi


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def _t(x, like=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if like is not None and t.dtype != like.dtype and t.is_complex():
        t = t.to(like.dtype)
    return t


def _front(state, dims, targets, is_matrix):
    """Tensor view with the targets (given order) leading: (t, r) or (t, r, t, r)."""
    dims = list(dims)
    targets = list(targets)
    rest = [i for i in range(len(dims)) if i not in targets]
    t = math.prod(dims[i] for i in targets)
    r = math.prod(dims[i] for i in rest)
    a = _np(state)
    if is_matrix:
        n = len(dims)
        perm = targets + rest
        a = a.reshape(dims + dims).transpose(perm + [n + p for p in perm]).reshape(t, r, t, r)
    else:
        a = a.reshape(dims).transpose(targets + rest).reshape(t, r)
    return a, t, r


def apply_vec(psi, op, dims, targets, out=None):
    return _t(oad.apply_operation_vector(tuple(dims), tuple(targets), _np(psi).reshape(-1, 1), _np(op)), psi)


def apply_mat(rho, op, dims, targets, out=None):
    n = math.prod(dims)
    return _t(oad.apply_operation_matrix(tuple(dims), tuple(targets), _np(rho).reshape(n, n), _np(op)), rho)


def kraus_mat(rho, ops, dims, targets, out=None):
    n = math.prod(dims)
    return _t(oad.apply_kraus_matrix(tuple(dims), tuple(targets), _np(rho).reshape(n, n), _np(ops)), rho)


def trace_out_mat(rho, dims, keep):
    n = math.prod(dims)
    if len(keep) == len(dims) and list(keep) == list(range(len(dims))):
        return rho
    a, t, r = _front(_np(rho).reshape(n, n), dims, keep, True)
    return _t(np.einsum("arbr->ab", a), rho)


def reduced_from_vec(psi, dims, keep):
    a, t, r = _front(psi, dims, keep, False)
    return _t(a @ a.conj().T, psi)


def probs(state, dims, targets, is_matrix):
    a, t, r = _front(state, dims, targets, is_matrix)
    if is_matrix:
        p = np.real(np.einsum("arar->a", a))
    else:
        p = (np.abs(a) ** 2).sum(axis=1)
    return torch.as_tensor(p / p.sum())


def draw(probs_t, key, cdtype):
    p = _np(probs_t).astype(np.float64)
    return torch.as_tensor(int(orng.choice(key, len(p), p)), dtype=torch.int64)


def collapse(state, outcome, probs_t, dims, targets, is_matrix):
    a, t, r = _front(state, dims, targets, is_matrix)
    o = int(outcome)
    if is_matrix:
        post = a[o, :, o, :]
        if probs_t is not None:
            post = post / (float(probs_t[o]) + 1e-18)
        return _t(post.reshape(r, r), state)
    post = a[o, :]
    if probs_t is not None:
        post = post / np.sqrt(float(probs_t[o]) + 1e-18)
    return _t(post.reshape(r, 1), state)


def povm_probs(rho_t, ops):
    rho, e = _np(rho_t), _np(ops)
    tr = np.array([np.real(np.trace(k @ rho @ k.conj().T)) for k in e])
    return torch.as_tensor(tr / tr.sum()), torch.as_tensor(tr)


def povm_select(ops, outcome, traces):
    o = int(outcome)
    return _t(_np(ops)[o] / np.sqrt(float(traces[o]) + 1e-18), ops)


def norm_stats(x):
    a = np.abs(_np(x)) ** 2
    return torch.tensor([a.sum(), a.max() if a.size else 0.0], dtype=torch.float64)


def find_label(psi):
    v = _np(psi).reshape(-1)
    hits = np.where(v == 1)[0]
    return torch.tensor([len(hits), int(hits[0]) if len(hits) else -1], dtype=torch.int64)


def occupancy(state, dims, is_matrix=False):
    a = _np(state)
    n = math.prod(dims)
    diag = np.real(np.diagonal(a.reshape(n, n))) if is_matrix else np.abs(a.reshape(-1)) ** 2
    w = diag.reshape(list(dims))
    levels = []
    for i in range(len(dims)):
        marg = np.abs(w).sum(axis=tuple(j for j in range(len(dims)) if j != i))
        nz = np.nonzero(marg)[0]
        levels.append(int(nz[-1]) if len(nz) else -1)
    biggest = float((np.abs(a) ** 2).max())
    return torch.tensor(levels, dtype=torch.int32), torch.tensor([float(diag.sum()), biggest],
                                                                 dtype=torch.float64)


def scale_(x, scale_dev):
    torch.view_as_real(x).mul_(float(scale_dev[0]))
    return x


def divide_(x, divisor_dev):
    torch.view_as_real(x).div_(float(divisor_dev[0]))  # component-wise, like the kernel
    return x


def permute(x, dims, perm, out=None):
    return _t(_np(x).reshape(list(dims)).transpose(list(perm)), x)


def resize_axis(x, dims, axis, new_dim, is_matrix):
    fn = olay.resize_fock_matrix if is_matrix else olay.resize_fock_vector
    n = math.prod(dims)
    a = _np(x).reshape(n, n) if is_matrix else _np(x).reshape(n, 1)
    return _t(fn(list(dims), axis, new_dim, a), x)


def kron(a, b):
    return _t(np.kron(_np(a), _np(b)), a)


def expand_vec(psi):
    return _t(olay.state_expand_vector(_np(psi)), psi)


def contract_mat(rho, tol=1e-6):
    r = _np(rho)
    purity = float(np.real(np.sum(np.abs(r) ** 2)))
    psi, ok = olay.state_contract_matrix(r, tol)
    if not ok:
        psi = np.zeros((r.shape[0], 1), dtype=r.dtype)
    return (_t(psi, rho), torch.tensor([purity, float(np.real(np.trace(r)))], dtype=torch.float64),
            torch.tensor(int(ok), dtype=torch.int32))


_PATCHED = ["apply_vec", "apply_mat", "kraus_mat", "trace_out_mat", "reduced_from_vec", "probs", "draw",
            "collapse", "povm_probs", "povm_select", "norm_stats", "find_label", "occupancy", "scale_", "divide_",
            "permute", "resize_axis", "kron", "expand_vec", "contract_mat"]


@contextlib.contextmanager
def host_backend():
    """Route core.ffi to the oracle and core._arrays to CPU tensors for the duration."""
    saved_ffi = {name: getattr(ffi, name) for name in _PATCHED}
    saved_a = {name: getattr(A, name) for name in ("require_cuda", "to_device", "is_host", "device")}
    g = globals()
    for name in _PATCHED:
        setattr(ffi, name, g[name])
    A.require_cuda = lambda: None
    A.device = lambda: "cpu"
    A.is_host = lambda x: not isinstance(x, torch.Tensor)
    A.to_device = lambda x: x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    try:
        yield
    finally:
        for name, fn in saved_ffi.items():
            setattr(ffi, name, fn)
        for name, fn in saved_a.items():
            setattr(A, name, fn)
