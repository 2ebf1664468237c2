"""
Tensor-level wrappers over the C ABI (include/pw_b200.h): each function takes
CUDA ``torch.Tensor`` buffers, passes raw device pointers + the current stream to
libpwb200 and returns freshly allocated outputs.  No arithmetic happens here.
"""

from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import numpy as np
import torch

from .. import _lib
from . import _arrays as A


def pin_operator(op: torch.Tensor) -> torch.Tensor:
    """Promise that ``op`` (a device operator) stays alive and unchanged: the library keeps its
    structure analysis per (operator, dims, targets) and later applies skip it.  The pin is released
    when the tensor is garbage collected.  Returns ``op``."""
    import weakref

    lib = _lib.load()
    ptr = op.data_ptr()
    _lib.check(lib.pw_op_pin(ptr), "pw_op_pin")
    fin = weakref.finalize(op, lib.pw_op_unpin, ptr)
    fin.atexit = False   # at interpreter exit the CUDA context may be gone already: nothing to free
    return op


_COMMON_CACHE: dict = {}


def _common(dims: Sequence[int], idx: Sequence[int]):
    """(ctypes dims, n, ctypes indices, count); cached per (dims, indices) pair of tuples -- the
    launch-bound configs spend their time on the host."""
    if type(dims) is tuple and type(idx) is tuple:
        hit = _COMMON_CACHE.get((dims, idx))
        if hit is None:
            hit = (_lib.i64_array(dims), len(dims), _lib.i32_array(idx), len(idx))
            if len(_COMMON_CACHE) < 4096:
                _COMMON_CACHE[(dims, idx)] = hit
        return hit
    return _lib.i64_array(dims), len(dims), _lib.i32_array(idx), len(idx)


def apply_vec(psi: torch.Tensor, op: torch.Tensor, dims, targets, out: Optional[torch.Tensor] = None):
    lib = _lib.load()
    out = torch.empty_like(psi) if out is None else out
    d, n, t, nt = _common(dims, targets)
    _lib.check(
        lib.pw_apply_vec(psi.data_ptr(), out.data_ptr(), op.data_ptr(), d, n, t, nt,
                         A.code(psi.dtype), A.stream_ptr()),
        "pw_apply_vec",
    )
    return out


def apply_mat(rho: torch.Tensor, op: torch.Tensor, dims, targets, out: Optional[torch.Tensor] = None):
    lib = _lib.load()
    out = torch.empty_like(rho) if out is None else out
    d, n, t, nt = _common(dims, targets)
    _lib.check(
        lib.pw_apply_mat(rho.data_ptr(), out.data_ptr(), op.data_ptr(), d, n, t, nt,
                         A.code(rho.dtype), A.stream_ptr()),
        "pw_apply_mat",
    )
    return out


def kraus_mat(rho: torch.Tensor, ops: torch.Tensor, dims, targets, out: Optional[torch.Tensor] = None):
    lib = _lib.load()
    out = torch.empty_like(rho) if out is None else out
    d, n, t, nt = _common(dims, targets)
    _lib.check(
        lib.pw_kraus_mat(rho.data_ptr(), out.data_ptr(), ops.data_ptr(), int(ops.shape[0]), d, n,
                         t, nt, A.code(rho.dtype), A.stream_ptr()),
        "pw_kraus_mat",
    )
    return out


def trace_out_mat(rho: torch.Tensor, dims, keep):
    lib = _lib.load()
    t = math.prod(dims[i] for i in keep)
    out = torch.empty((t, t), dtype=rho.dtype, device=rho.device)
    d, n, k, nk = _common(dims, keep)
    _lib.check(
        lib.pw_trace_out_mat(rho.data_ptr(), out.data_ptr(), d, n, k, nk, A.code(rho.dtype),
                             A.stream_ptr()),
        "pw_trace_out_mat",
    )
    return out


def reduced_from_vec(psi: torch.Tensor, dims, keep):
    lib = _lib.load()
    t = math.prod(dims[i] for i in keep)
    out = torch.empty((t, t), dtype=psi.dtype, device=psi.device)
    d, n, k, nk = _common(dims, keep)
    _lib.check(
        lib.pw_reduced_from_vec(psi.data_ptr(), out.data_ptr(), d, n, k, nk, A.code(psi.dtype),
                                A.stream_ptr()),
        "pw_reduced_from_vec",
    )
    return out


def cross_reduced(x: torch.Tensor, y: torch.Tensor, axes, keep):
    """out[a, b] = sum_rest x[(a, rest)] conj(y[(b, rest)]) (operator cotangent of a linear apply)."""
    lib = _lib.load()
    t = math.prod(axes[i] for i in keep)
    out = torch.empty((t, t), dtype=x.dtype, device=x.device)
    _lib.check(
        lib.pw_cross_reduced(x.data_ptr(), y.data_ptr(), out.data_ptr(), _lib.i64_array(axes), len(axes),
                             _lib.i32_array(keep), len(keep), A.code(x.dtype), A.stream_ptr()),
        "pw_cross_reduced",
    )
    return out


def probs(state: torch.Tensor, dims, targets, is_matrix: bool):
    lib = _lib.load()
    t = math.prod(dims[i] for i in targets)
    out = torch.empty((t,), dtype=A.real_dtype(state.dtype), device=state.device)
    d, n, tg, nt = _common(dims, targets)
    fn = lib.pw_probs_mat if is_matrix else lib.pw_probs_vec
    _lib.check(
        fn(state.data_ptr(), out.data_ptr(), d, n, tg, nt, A.code(state.dtype), A.stream_ptr()),
        "pw_probs_mat" if is_matrix else "pw_probs_vec",
    )
    return out


def draw(probs_t: torch.Tensor, key, cdtype: torch.dtype):
    lib = _lib.load()
    k0, k1 = A.key_words(key)
    outcome = torch.empty((), dtype=torch.int64, device=probs_t.device)
    _lib.check(
        lib.pw_draw(probs_t.data_ptr(), int(probs_t.numel()), k0, k1, A.code(cdtype),
                    outcome.data_ptr(), A.stream_ptr()),
        "pw_draw",
    )
    return outcome


def collapse(state: torch.Tensor, outcome: torch.Tensor, probs_t: torch.Tensor, dims, targets,
             is_matrix: bool):
    lib = _lib.load()
    r = math.prod(dims) // math.prod(dims[i] for i in targets)
    shape = (r, r) if is_matrix else (r, 1)
    post = torch.empty(shape, dtype=state.dtype, device=state.device)
    d, n, tg, nt = _common(dims, targets)
    fn = lib.pw_collapse_mat if is_matrix else lib.pw_collapse_vec
    _lib.check(
        fn(state.data_ptr(), post.data_ptr(), outcome.data_ptr(),
           probs_t.data_ptr() if probs_t is not None else None, d, n, tg,
           nt, A.code(state.dtype), A.stream_ptr()),
        "pw_collapse_mat" if is_matrix else "pw_collapse_vec",
    )
    return post


def povm_probs(rho_t: torch.Tensor, ops: torch.Tensor):
    lib = _lib.load()
    K, t = int(ops.shape[0]), int(rho_t.shape[0])
    rdt = A.real_dtype(rho_t.dtype)
    p = torch.empty((K,), dtype=rdt, device=rho_t.device)
    tr = torch.empty((K,), dtype=rdt, device=rho_t.device)
    _lib.check(
        lib.pw_povm_probs(rho_t.data_ptr(), ops.data_ptr(), K, t, p.data_ptr(), tr.data_ptr(),
                          A.code(rho_t.dtype), A.stream_ptr()),
        "pw_povm_probs",
    )
    return p, tr


def povm_select(ops: torch.Tensor, outcome: torch.Tensor, traces: torch.Tensor):
    lib = _lib.load()
    t = int(ops.shape[-1])
    out = torch.empty((t, t), dtype=ops.dtype, device=ops.device)
    _lib.check(
        lib.pw_povm_select(ops.data_ptr(), t, outcome.data_ptr(), traces.data_ptr(),
                           out.data_ptr(), A.code(ops.dtype), A.stream_ptr()),
        "pw_povm_select",
    )
    return out


def norm_stats(x: torch.Tensor):
    """Device tensor [sum |x|^2, max |x|^2] (float64), no host sync."""
    lib = _lib.load()
    stats = torch.empty((2,), dtype=torch.float64, device=x.device)
    _lib.check(
        lib.pw_norm_stats(x.data_ptr(), int(x.numel()), A.code(x.dtype), stats.data_ptr(),
                          A.stream_ptr()),
        "pw_norm_stats",
    )
    return stats


def find_label(psi: torch.Tensor):
    """Device int64 [count of entries == 1+0j, index of the first] -- no host sync."""
    lib = _lib.load()
    out = torch.empty((2,), dtype=torch.int64, device=psi.device)
    _lib.check(
        lib.pw_find_label(psi.data_ptr(), int(psi.numel()), A.code(psi.dtype), out.data_ptr(),
                          A.stream_ptr()),
        "pw_find_label",
    )
    return out


def occupancy(state: torch.Tensor, dims, is_matrix: bool = False):
    """One pass: (levels int32[n], stats float64[2]) -- highest non-zero basis index of every
    subsystem, and [sum |psi|^2 or Tr rho, max |x|^2].  Device tensors, no host sync."""
    lib = _lib.load()
    d = _lib.i64_array(dims)
    levels = torch.empty((len(dims),), dtype=torch.int32, device=state.device)
    stats = torch.empty((2,), dtype=torch.float64, device=state.device)
    _lib.check(
        lib.pw_occupancy(state.data_ptr(), d, len(dims), int(bool(is_matrix)),
                         A.code(state.dtype), levels.data_ptr(), stats.data_ptr(),
                         A.stream_ptr()),
        "pw_occupancy",
    )
    return levels, stats


def scale_(x: torch.Tensor, scale_dev: torch.Tensor):
    lib = _lib.load()
    _lib.check(
        lib.pw_scale(x.data_ptr(), int(x.numel()), scale_dev.data_ptr(), A.code(x.dtype),
                     A.stream_ptr()),
        "pw_scale",
    )
    return x


def divide_(x: torch.Tensor, divisor_dev: torch.Tensor):
    lib = _lib.load()
    _lib.check(
        lib.pw_divide(x.data_ptr(), int(x.numel()), divisor_dev.data_ptr(), A.code(x.dtype),
                      A.stream_ptr()),
        "pw_divide",
    )
    return x


def permute(x: torch.Tensor, dims, perm, out: Optional[torch.Tensor] = None):
    """out = transpose(x.reshape(dims), perm) made contiguous (flat tensor returned)."""
    lib = _lib.load()
    out = torch.empty_like(x) if out is None else out
    d, n, p, _ = _common(dims, perm)
    _lib.check(
        lib.pw_permute(x.data_ptr(), out.data_ptr(), d, n, p, A.code(x.dtype), A.stream_ptr()),
        "pw_permute",
    )
    return out


def apply_vec_pair(psi: torch.Tensor, op_a: torch.Tensor, op_b: torch.Tensor, dims, axis_a: int,
                   axis_b: int, out: Optional[torch.Tensor] = None):
    """(A on axis_a)(B on axis_b) psi in one pass; returns None when the shape is unsupported."""
    lib = _lib.load()
    out = torch.empty_like(psi) if out is None else out
    rc = lib.pw_apply_vec_pair(psi.data_ptr(), out.data_ptr(), op_a.data_ptr(), op_b.data_ptr(),
                               _lib.i64_array(dims), len(dims), int(axis_a), int(axis_b),
                               A.code(psi.dtype), A.stream_ptr())
    if rc == 2:
        return None
    _lib.check(rc, "pw_apply_vec_pair")
    return out


def apply_vec_multi(psi: torch.Tensor, ops, dims, axes, out: Optional[torch.Tensor] = None):
    """prod_i (ops[i] on axes[i]) psi in ONE pass (2..4 operators of one dimension on different
    axes); returns None when the shape is unsupported."""
    lib = _lib.load()
    out = torch.empty_like(psi) if out is None else out
    arr = (C.c_void_p * len(ops))(*[o.data_ptr() for o in ops])
    rc = lib.pw_apply_vec_multi(psi.data_ptr(), out.data_ptr(), arr, _lib.i64_array(dims), len(dims),
                                _lib.i32_array(axes), len(axes), A.code(psi.dtype), A.stream_ptr())
    if rc == 2:
        return None
    _lib.check(rc, "pw_apply_vec_multi")
    return out


def circuit_device(state: torch.Tensor, scratch: Optional[torch.Tensor], is_matrix: bool, dims,
                   ops_list, fuse: bool = True):
    """Device-resident op list (pw_circuit_device): ops_list = [(kind, targets, device tensor)].
    Returns (result, other) -- the two buffers, result first."""
    lib = _lib.load()
    arr = (_lib.pw_host_op * max(len(ops_list), 1))()
    for i, (kind, targets, op) in enumerate(ops_list):
        arr[i].kind = int(kind)
        arr[i].nt = len(targets)
        for k, tg in enumerate(targets):
            arr[i].targets[k] = int(tg)
        arr[i].K = int(op.shape[0]) if kind == 2 else 1
        arr[i].op = op.data_ptr()
    flag = C.c_int(0)
    _lib.check(
        lib.pw_circuit_device(state.data_ptr(), scratch.data_ptr() if scratch is not None else None,
                              int(is_matrix), _lib.i64_array(dims), len(dims), arr, len(ops_list),
                              A.code(state.dtype), int(fuse), A.stream_ptr(), C.byref(flag)),
        "pw_circuit_device",
    )
    return (scratch, state) if flag.value else (state, scratch)


def evolve_vec(psi: torch.Tensor, op: torch.Tensor, dims, nsteps: int, keep=None):
    """psi <- op^nsteps psi in one kernel (small states); returns (psi_out, reduced[nsteps,tk,tk]
    or None)."""
    lib = _lib.load()
    out = torch.empty_like(psi)
    red = None
    keep = list(keep) if keep is not None else []
    if keep:
        tk = math.prod(dims[i] for i in keep)
        red = torch.empty((nsteps, tk, tk), dtype=psi.dtype, device=psi.device)
    _lib.check(
        lib.pw_evolve_vec(psi.data_ptr(), out.data_ptr(), op.data_ptr(), _lib.i64_array(dims), len(dims),
                          _lib.i32_array(keep), len(keep), int(nsteps),
                          red.data_ptr() if red is not None else None, A.code(psi.dtype),
                          A.stream_ptr()),
        "pw_evolve_vec",
    )
    return out, red


def resize_axis(x: torch.Tensor, dims, axis: int, new_dim: int, is_matrix: bool):
    """Pad / truncate one subsystem (both row and column copies for a density matrix)."""
    lib = _lib.load()
    nd = list(dims)
    nd[axis] = int(new_dim)
    n_new = math.prod(nd)
    shape = (n_new, n_new) if is_matrix else (n_new, 1)
    out = torch.empty(shape, dtype=x.dtype, device=x.device)
    _lib.check(
        lib.pw_resize_axis(x.data_ptr(), out.data_ptr(), _lib.i64_array(dims), len(dims), int(axis),
                           int(new_dim), int(is_matrix), A.code(x.dtype), A.stream_ptr()),
        "pw_resize_axis",
    )
    return out


def kron(a: torch.Tensor, b: torch.Tensor):
    lib = _lib.load()
    ra, ca = a.shape
    rb, cb = b.shape
    out = torch.empty((ra * rb, ca * cb), dtype=a.dtype, device=a.device)
    _lib.check(
        lib.pw_kron(a.data_ptr(), b.data_ptr(), out.data_ptr(), ra, ca, rb, cb, A.code(a.dtype),
                    A.stream_ptr()),
        "pw_kron",
    )
    return out


def expand_vec(psi: torch.Tensor):
    lib = _lib.load()
    n = psi.numel()
    out = torch.empty((n, n), dtype=psi.dtype, device=psi.device)
    _lib.check(lib.pw_expand_vec(psi.data_ptr(), out.data_ptr(), n, A.code(psi.dtype), A.stream_ptr()),
               "pw_expand_vec")
    return out


def contract_mat(rho: torch.Tensor, tol: float = 1e-6):
    """(psi (n,1), info [Tr rho^2, Tr rho], flag) -- all device tensors, no host sync."""
    lib = _lib.load()
    n = int(rho.shape[0])
    psi = torch.zeros((n, 1), dtype=rho.dtype, device=rho.device)
    info = torch.empty((2,), dtype=torch.float64, device=rho.device)
    flag = torch.empty((), dtype=torch.int32, device=rho.device)
    _lib.check(
        lib.pw_contract_mat(rho.data_ptr(), n, float(tol), psi.data_ptr(), info.data_ptr(),
                            flag.data_ptr(), A.code(rho.dtype), A.stream_ptr()),
        "pw_contract_mat",
    )
    return psi, info, flag


# ---- blocks of a sharded state (multi-GPU layer, photon_weave_b200/sharded.py) ---------------


def _view(pairs):
    """[(extent, stride)] -> two ctypes int64 arrays + count."""
    ext = [int(e) for e, _ in pairs]
    st = [int(s) for _, s in pairs]
    return _lib.i64_array(ext), _lib.i64_array(st), len(pairs)


def _esize(dtype: torch.dtype) -> int:
    return 16 if dtype == torch.complex128 else 8


def kraus_axes(x: torch.Tensor, ops: torch.Tensor, axes, row_targets, col_targets,
               out: Optional[torch.Tensor] = None):
    """out = sum_k (K_k on row_targets)(conj K_k on col_targets) x, x a tensor over ``axes``."""
    lib = _lib.load()
    out = torch.empty_like(x) if out is None else out
    _lib.check(
        lib.pw_kraus_axes(x.data_ptr(), out.data_ptr(), ops.data_ptr(), int(ops.shape[0]),
                          _lib.i64_array(axes), len(axes), _lib.i32_array(row_targets), len(row_targets),
                          _lib.i32_array(col_targets), len(col_targets), A.code(x.dtype), A.stream_ptr()),
        "pw_kraus_axes",
    )
    return out


def sums_strided(x: torch.Tensor, offset: int, rest, tgt, real_part: bool):
    """Un-normalised float64 sums over the ``rest`` view per joint index of the ``tgt`` view."""
    lib = _lib.load()
    t = math.prod(e for e, _ in tgt) if tgt else 1
    out = torch.empty((t,), dtype=torch.float64, device=x.device)
    re, rs, nr = _view(rest)
    te, ts, nt = _view(tgt)
    _lib.check(
        lib.pw_sums_strided(x.data_ptr() + int(offset) * _esize(x.dtype), re, rs, nr, te, ts, nt,
                            int(bool(real_part)), out.data_ptr(), A.code(x.dtype), A.stream_ptr()),
        "pw_sums_strided",
    )
    return out


def probs_from_sums(sums: torch.Tensor, cdtype: torch.dtype):
    lib = _lib.load()
    out = torch.empty((sums.numel(),), dtype=A.real_dtype(cdtype), device=sums.device)
    _lib.check(lib.pw_probs_from_sums(sums.data_ptr(), int(sums.numel()), out.data_ptr(), A.code(cdtype),
                                      A.stream_ptr()), "pw_probs_from_sums")
    return out


def trace_strided(x: torch.Tensor, offset: int, rest, rows, cols):
    lib = _lib.load()
    tr = math.prod(e for e, _ in rows) if rows else 1
    tc = math.prod(e for e, _ in cols) if cols else 1
    out = torch.empty((tr, tc), dtype=x.dtype, device=x.device)
    re, rs, nr = _view(rest)
    we, ws, nw = _view(rows)
    ce, cs, nc = _view(cols)
    _lib.check(
        lib.pw_trace_strided(x.data_ptr() + int(offset) * _esize(x.dtype), re, rs, nr, we, ws, nw,
                             ce, cs, nc, out.data_ptr(), A.code(x.dtype), A.stream_ptr()),
        "pw_trace_strided",
    )
    return out


def gather_strided(x: torch.Tensor, offset: int, view, probs_t, outcome, mode: int):
    """Contiguous copy of a strided view, divided by 1 / sqrt(p_o + 1e-18) / (p_o + 1e-18)."""
    lib = _lib.load()
    n = math.prod(e for e, _ in view) if view else 1
    out = torch.empty((n,), dtype=x.dtype, device=x.device)
    ve, vs, nv = _view(view)
    _lib.check(
        lib.pw_gather_strided(x.data_ptr() + int(offset) * _esize(x.dtype), ve, vs, nv, out.data_ptr(),
                              probs_t.data_ptr() if probs_t is not None else None,
                              outcome.data_ptr() if outcome is not None else None, int(mode),
                              A.code(x.dtype), A.stream_ptr()),
        "pw_gather_strided",
    )
    return out


def rng_split(key, num: int = 2) -> np.ndarray:
    lib = _lib.load()
    k0, k1 = A.key_words(key)
    kin = (C.c_uint32 * 2)(k0, k1)
    out = (C.c_uint32 * (2 * num))()
    _lib.check(lib.pw_rng_split(kin, num, out), "pw_rng_split")
    return np.frombuffer(out, dtype=np.uint32).reshape(num, 2).copy()


def _host_ops(ops_list, cdt):
    arr = (_lib.pw_host_op * max(len(ops_list), 1))()
    keep = []
    for i, (kind, targets, op) in enumerate(ops_list):
        op = np.ascontiguousarray(op, dtype=cdt)
        keep.append(op)
        arr[i].kind = int(kind)
        arr[i].nt = len(targets)
        for k, tg in enumerate(targets):
            arr[i].targets[k] = int(tg)
        arr[i].K = int(op.shape[0]) if kind == 2 else 1
        arr[i].op = op.ctypes.data
    return arr, keep


class HostPipeline:
    """
    Streaming host-buffer path (pw_pipeline_*): ``submit`` enqueues upload + operation list +
    download of one state and returns a ticket, ``wait`` blocks until that result is in its
    output buffer.  Uploads, kernels and downloads of consecutive jobs overlap (PCIe full
    duplex).  Host buffers should be pinned and must not be shared by jobs in flight.
    """

    def __init__(self, dims, is_matrix: bool = False, dtype=np.complex128, device: int = 0,
                 buffers: int = 4):
        self._lib = _lib.load()
        self.dims = tuple(int(d) for d in dims)
        self.is_matrix = bool(is_matrix)
        self.cdt = np.complex128 if np.dtype(dtype) == np.complex128 else np.complex64
        self._h = C.c_void_p()
        _lib.check(
            self._lib.pw_pipeline_create(int(is_matrix), _lib.i64_array(self.dims), len(self.dims),
                                         _lib.PW_C128 if self.cdt == np.complex128 else _lib.PW_C64,
                                         int(device), int(buffers), C.byref(self._h)),
            "pw_pipeline_create")
        self._keep = {}

    @property
    def buffers(self) -> int:
        return int(self._lib.pw_pipeline_buffers(self._h))

    def submit(self, h_state: np.ndarray, out: np.ndarray, ops_list) -> int:
        assert h_state.dtype == self.cdt and out.dtype == self.cdt
        assert h_state.flags.c_contiguous and out.flags.c_contiguous
        arr, keep = _host_ops(ops_list, self.cdt)
        ticket = C.c_int64(-1)
        _lib.check(self._lib.pw_pipeline_submit(self._h, h_state.ctypes.data, out.ctypes.data, arr,
                                                len(ops_list), C.byref(ticket)),
                   "pw_pipeline_submit")
        self._keep[ticket.value] = (keep, h_state, out)
        return int(ticket.value)

    def wait(self, ticket: int) -> np.ndarray:
        _lib.check(self._lib.pw_pipeline_wait(self._h, int(ticket)), "pw_pipeline_wait")
        _, _, out = self._keep.pop(int(ticket), (None, None, None))
        return out

    def close(self) -> None:
        if self._h:
            self._lib.pw_pipeline_destroy(self._h)
            self._h = C.c_void_p()
            self._keep.clear()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def circuit_host(h_state: np.ndarray, is_matrix: bool, dims, ops_list, device: int = 0,
                 out: Optional[np.ndarray] = None) -> np.ndarray:
    """
    End-to-end host-buffer path (pw_circuit_host): ops_list holds
    (kind, targets, operator ndarray) with kind 0 vector-op, 1 matrix-op, 2 Kraus.
    """
    lib = _lib.load()
    cdt = np.complex128 if h_state.dtype == np.complex128 else np.complex64
    h_state = np.ascontiguousarray(h_state, dtype=cdt)
    out = np.empty_like(h_state) if out is None else out
    arr = (_lib.pw_host_op * max(len(ops_list), 1))()
    keep = []
    for i, (kind, targets, op) in enumerate(ops_list):
        op = np.ascontiguousarray(op, dtype=cdt)
        keep.append(op)
        arr[i].kind = int(kind)
        arr[i].nt = len(targets)
        for k, tg in enumerate(targets):
            arr[i].targets[k] = int(tg)
        arr[i].K = int(op.shape[0]) if kind == 2 else 1
        arr[i].op = op.ctypes.data
    _lib.check(
        lib.pw_circuit_host(h_state.ctypes.data, out.ctypes.data, int(is_matrix),
                            _lib.i64_array(dims), len(dims), arr, len(ops_list),
                            _lib.PW_C128 if cdt == np.complex128 else _lib.PW_C64, device),
        "pw_circuit_host",
    )
    return out
