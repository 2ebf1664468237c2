"""
ctypes binding of ``libpwb200.so`` (C ABI declared in ``include/pw_b200.h``).

There is NO CPU fallback: if the shared library is missing or a call fails, the
product path raises.  Build it with ``python -c "import __graft_entry__ as g;
g.build()"`` (or ``make -C photon_weave_b200/csrc``).
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpwb200.so")

PW_C64, PW_C128 = 0, 1
PW_OK = 0
PW_MAX_TARGETS = 8


class PwError(RuntimeError):
    """A libpwb200 entry point returned a non-zero status."""


class pw_host_op(C.Structure):
    _fields_ = [
        ("kind", C.c_int),
        ("nt", C.c_int),
        ("targets", C.c_int32 * PW_MAX_TARGETS),
        ("K", C.c_int),
        ("op", C.c_void_p),
    ]


_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/pw_b200.h one to one
SIGNATURES = {
    "pw_version": (C.c_int, []),
    "pw_status_string": (C.c_char_p, [C.c_int]),
    "pw_last_cuda_error": (C.c_int, []),
    "pw_launch_count": (C.c_uint64, []),
    "pw_path_counters": (None, [C.POINTER(C.c_uint64)]),
    "pw_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "pw_probe_arith": (C.c_int, [C.c_int, C.c_int, _vp, C.POINTER(C.c_double), _vp]),
    "pw_op_pin": (C.c_int, [_vp]),
    "pw_op_unpin": (C.c_int, [_vp]),
    "pw_plan_cache_size": (C.c_int, []),
    "pw_apply_vec": (C.c_int, [_vp, _vp, _vp, _i64p, C.c_int, _i32p, C.c_int, C.c_int, _vp]),
    "pw_apply_mat": (C.c_int, [_vp, _vp, _vp, _i64p, C.c_int, _i32p, C.c_int, C.c_int, _vp]),
    "pw_kraus_mat": (
        C.c_int,
        [_vp, _vp, _vp, C.c_int, _i64p, C.c_int, _i32p, C.c_int, C.c_int, _vp],
    ),
    "pw_trace_out_mat": (C.c_int, [_vp, _vp, _i64p, C.c_int, _i32p, C.c_int, C.c_int, _vp]),
    "pw_reduced_from_vec": (C.c_int, [_vp, _vp, _i64p, C.c_int, _i32p, C.c_int, C.c_int, _vp]),
    "pw_cross_reduced": (C.c_int, [_vp, _vp, _vp, _i64p, C.c_int, _i32p, C.c_int, C.c_int, _vp]),
    "pw_probs_vec": (C.c_int, [_vp, _vp, _i64p, C.c_int, _i32p, C.c_int, C.c_int, _vp]),
    "pw_probs_mat": (C.c_int, [_vp, _vp, _i64p, C.c_int, _i32p, C.c_int, C.c_int, _vp]),
    "pw_draw": (C.c_int, [_vp, C.c_int64, C.c_uint32, C.c_uint32, C.c_int, _vp, _vp]),
    "pw_draw_devkey": (C.c_int, [_vp, C.c_int64, _vp, C.c_int, _vp, _vp]),
    "pw_collapse_vec": (
        C.c_int,
        [_vp, _vp, _vp, _vp, _i64p, C.c_int, _i32p, C.c_int, C.c_int, _vp],
    ),
    "pw_collapse_mat": (
        C.c_int,
        [_vp, _vp, _vp, _vp, _i64p, C.c_int, _i32p, C.c_int, C.c_int, _vp],
    ),
    "pw_povm_probs": (C.c_int, [_vp, _vp, C.c_int, C.c_int64, _vp, _vp, C.c_int, _vp]),
    "pw_povm_select": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, C.c_int, _vp]),
    "pw_norm_stats": (C.c_int, [_vp, C.c_int64, C.c_int, _vp, _vp]),
    "pw_find_label": (C.c_int, [_vp, C.c_int64, C.c_int, _vp, _vp]),
    "pw_occupancy": (C.c_int, [_vp, _i64p, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "pw_scale": (C.c_int, [_vp, C.c_int64, _vp, C.c_int, _vp]),
    "pw_divide": (C.c_int, [_vp, C.c_int64, _vp, C.c_int, _vp]),
    "pw_permute": (C.c_int, [_vp, _vp, _i64p, C.c_int, _i32p, C.c_int, _vp]),
    "pw_resize_axis": (C.c_int, [_vp, _vp, _i64p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, _vp]),
    "pw_kron": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, _vp]),
    "pw_expand_vec": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, _vp]),
    "pw_contract_mat": (C.c_int, [_vp, C.c_int64, C.c_double, _vp, _vp, _vp, C.c_int, _vp]),
    "pw_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp), C.POINTER(C.c_ubyte)]),
    "pw_peer_open": (C.c_int, [C.POINTER(C.c_ubyte), C.POINTER(_vp)]),
    "pw_peer_close": (C.c_int, [_vp]),
    "pw_peer_free": (C.c_int, [_vp]),
    "pw_peer_copy": (C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    "pw_gather_chunks": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int64, C.c_int, _vp, C.c_int, _vp]),
    "pw_apply_vec_gather": (
        C.c_int,
        [C.POINTER(_vp), C.c_int, C.c_int64, C.c_int, _vp, _vp, _i64p, C.c_int, _i32p, C.c_int, C.c_int, _vp],
    ),
    "pw_kraus_axes": (
        C.c_int,
        [_vp, _vp, _vp, C.c_int, _i64p, C.c_int, _i32p, C.c_int, _i32p, C.c_int, C.c_int, _vp],
    ),
    "pw_sums_strided": (
        C.c_int,
        [_vp, _i64p, _i64p, C.c_int, _i64p, _i64p, C.c_int, C.c_int, _vp, C.c_int, _vp],
    ),
    "pw_probs_from_sums": (C.c_int, [_vp, C.c_int64, _vp, C.c_int, _vp]),
    "pw_trace_strided": (
        C.c_int,
        [_vp, _i64p, _i64p, C.c_int, _i64p, _i64p, C.c_int, _i64p, _i64p, C.c_int, _vp, C.c_int, _vp],
    ),
    "pw_gather_strided": (
        C.c_int,
        [_vp, _i64p, _i64p, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp],
    ),
    "pw_rng_split": (C.c_int, [C.POINTER(C.c_uint32), C.c_int, C.POINTER(C.c_uint32)]),
    "pw_apply_vec_pair": (C.c_int, [_vp, _vp, _vp, _vp, _i64p, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "pw_apply_vec_multi": (C.c_int, [_vp, _vp, C.POINTER(_vp), _i64p, C.c_int, _i32p, C.c_int, C.c_int, _vp]),
    "pw_evolve_vec": (
        C.c_int,
        [_vp, _vp, _vp, _i64p, C.c_int, _i32p, C.c_int, C.c_int, _vp, C.c_int, _vp],
    ),
    "pw_circuit_device": (
        C.c_int,
        [_vp, _vp, C.c_int, _i64p, C.c_int, C.POINTER(pw_host_op), C.c_int, C.c_int, C.c_int, _vp,
         C.POINTER(C.c_int)],
    ),
    "pw_circuit_host": (
        C.c_int,
        [_vp, _vp, C.c_int, _i64p, C.c_int, C.POINTER(pw_host_op), C.c_int, C.c_int, C.c_int],
    ),
    "pw_pipeline_create": (C.c_int, [C.c_int, _i64p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_void_p)]),
    "pw_pipeline_buffers": (C.c_int, [_vp]),
    "pw_pipeline_submit": (C.c_int, [_vp, _vp, _vp, C.POINTER(pw_host_op), C.c_int,
                                     C.POINTER(C.c_int64)]),
    "pw_pipeline_wait": (C.c_int, [_vp, C.c_int64]),
    "pw_pipeline_destroy": (C.c_int, [_vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load libpwb200.so (once) and declare every prototype.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA library is not built and there is no CPU "
            "fallback. Run `python -c 'import __graft_entry__ as g; g.build()'`."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != PW_OK:
        lib = load()
        msg = lib.pw_status_string(status).decode()
        extra = f" (cudaError {lib.pw_last_cuda_error()})" if status == 3 else ""
        raise PwError(f"{what}: {msg}{extra}")


_I64_CACHE: dict = {}
_I32_CACHE: dict = {}


def i64_array(vals: Sequence[int]):
    """ctypes int64 array of ``vals`` (read-only to the library: cached per value tuple -- the
    per-call latency of the small configs is host-bound)."""
    key = vals if type(vals) is tuple else tuple(int(v) for v in vals)
    arr = _I64_CACHE.get(key)
    if arr is None:
        arr = (C.c_int64 * max(len(key), 1))(*[int(v) for v in key])
        if len(_I64_CACHE) < 4096:
            _I64_CACHE[key] = arr
    return arr


def i32_array(vals: Sequence[int]):
    key = vals if type(vals) is tuple else tuple(int(v) for v in vals)
    arr = _I32_CACHE.get(key)
    if arr is None:
        arr = (C.c_int32 * max(len(key), 1))(*[int(v) for v in key])
        if len(_I32_CACHE) < 4096:
            _I32_CACHE[key] = arr
    return arr


def path_counters() -> dict:
    """{'register': n, 'generic': n, 'tiled': n, 'tiled_rejected': n} since library load."""
    buf = (C.c_uint64 * 5)()
    load().pw_path_counters(buf)
    names = ("register", "generic", "tiled", "tiled_rejected", "blocks_attempted")
    return dict(zip(names, [int(v) for v in buf]))


def set_option(name: str, value: int) -> None:
    check(load().pw_set_option(name.encode(), int(value)), f"pw_set_option({name})")
