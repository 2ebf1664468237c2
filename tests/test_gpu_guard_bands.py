"""Own bounds check of the kernel paths (compute-sanitizer is not available on the GPU pool):
inputs and outputs are views in the MIDDLE of larger buffers filled with a sentinel; after every
call the bands on both sides of the output must be untouched, the input must be unchanged, and
the result must equal the oracle's.  Shapes are chosen to hit every kernel family: register,
complex64 fiber pairs, diagonal, block-structured, staged (innermost target), dense and two-sided
tensor-core kernels, the fused tile kernel, ragged tails included."""
import math

import numpy as np
import pytest

from oracle import adapters_np as oad
from oracle import operators as ops

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

GUARD = 4099  # elements on each side (odd on purpose: the views are not 16/32-byte aligned for c64)
TOL = {np.complex128: 1e-12, np.complex64: 1e-5}


def _rel(a, b):
    return float(np.linalg.norm(a.reshape(-1) - b.reshape(-1)) / max(np.linalg.norm(b), 1e-300))


def _banded(n, tdtype, fill):
    buf = torch.full((n + 2 * GUARD,), fill, dtype=tdtype, device="cuda")
    return buf, buf[GUARD:GUARD + n]


def _check_bands(buf, n, fill, what):
    head, tail = buf[:GUARD], buf[GUARD + n:]
    assert bool((head == fill).all()) and bool((tail == fill).all()), f"{what}: wrote outside the output"


VEC_CASES = [
    ((6, 5, 4), (0,), "dense"), ((3, 6, 7, 2), (1,), "dense"), ((8, 3, 10), (2,), "dense"),
    ((4, 7, 9), (1,), "phase"), ((6, 6, 6, 6, 36), (1,), "phase"),
    ((5, 5, 6), (0, 1), "bs"), ((4, 6, 6, 14), (1, 2), "bs"), ((7, 6, 6), (1, 2), "bs"), ((6, 11, 6), (2, 0), "bs"),
    ((3, 36, 40), (1,), "dense"), ((33, 20), (1,), "dense"), ((5, 9, 11, 13), (1, 2), "dense"),
    ((2, 70, 5), (1,), "dense"), ((9, 128), (1,), "dense"),
]
MAT_CASES = [
    ((4, 3), (0,), "dense"), ((6, 5), (1,), "dense"), ((8, 3, 2), (0,), "dense"), ((2, 3, 7), (2,), "dense"),
    ((5, 5, 3), (0, 1), "bs"), ((3, 4, 4), (1, 2), "bs"), ((4, 3, 4), (2, 0), "bs"),
    ((6, 4), (0,), "loss"), ((3, 8), (1,), "loss"), ((5, 4, 2), (1,), "loss"),
]


def _operator(kind, dims, tg, cd):
    t = math.prod(dims[i] for i in tg)
    if kind == "bs":
        op = ops.beam_splitter(dims[tg[0]], dims[tg[1]], 0.37)
    elif kind == "phase":
        op = ops.phase_operator(t, 0.7)
    elif kind == "loss":
        op = ops.loss_channel(t, 0.2)
    else:
        op = ops.random_unitary(t, 3)
    return np.ascontiguousarray(op).astype(cd)


@pytest.mark.parametrize("cd", [np.complex128, np.complex64])
def test_vector_kernels_stay_inside_their_buffers(cd):
    from photon_weave_b200 import _lib
    from photon_weave_b200.core import ffi

    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    td = torch.complex128 if cd == np.complex128 else torch.complex64
    fill = complex(7.25, -3.5)
    _lib.set_option("blocks_min_elems", 0)
    try:
        for dims, tg, kind in VEC_CASES:
            N = math.prod(dims)
            op = _operator(kind, dims, tg, cd)
            psi = ops.random_state(N, 5, cd).reshape(-1)
            ibuf, iview = _banded(N, td, fill)
            obuf, oview = _banded(N, td, fill)
            iview.copy_(torch.as_tensor(psi).cuda())
            ffi.apply_vec(iview, torch.as_tensor(op).cuda(), dims, tg, out=oview)
            torch.cuda.synchronize()
            _check_bands(obuf, N, fill, (dims, tg, kind))
            _check_bands(ibuf, N, fill, (dims, tg, kind, "input"))
            assert np.array_equal(iview.cpu().numpy(), psi), (dims, tg, kind, "input modified")
            ref = oad.apply_operation_vector(dims, tg, psi.astype(np.complex128).reshape(-1, 1),
                                             op.astype(np.complex128), True)
            assert _rel(oview.cpu().numpy(), ref) < TOL[cd], (dims, tg, kind)
        # fused tile kernel: three operators on three axes, ragged fiber counts
        for dims, axes in [((6, 6, 6, 7), (0, 1, 2)), ((5, 3, 5, 5, 11), (0, 2, 3)), ((4, 4, 9, 4), (0, 1, 3))]:
            N = math.prod(dims)
            d = dims[axes[0]]
            us = [np.ascontiguousarray(ops.random_unitary(d, 20 + k)).astype(cd) for k in range(3)]
            psi = ops.random_state(N, 6, cd).reshape(-1)
            ibuf, iview = _banded(N, td, fill)
            obuf, oview = _banded(N, td, fill)
            iview.copy_(torch.as_tensor(psi).cuda())
            res = ffi.apply_vec_multi(iview, [torch.as_tensor(u).cuda() for u in us], dims, axes, out=oview)
            torch.cuda.synchronize()
            _check_bands(obuf, N, fill, (dims, axes, "multi"))
            if res is None:
                continue   # shape declined: nothing may have been written
            ref = psi.astype(np.complex128).reshape(-1, 1)
            for u, a in zip(us, axes):
                ref = oad.apply_operation_vector(dims, (a,), ref, u.astype(np.complex128), True)
            assert _rel(oview.cpu().numpy(), ref) < TOL[cd], (dims, axes)
    finally:
        _lib.set_option("blocks_min_elems", 1 << 18)


@pytest.mark.parametrize("cd", [np.complex128, np.complex64])
def test_matrix_kernels_stay_inside_their_buffers(cd):
    from photon_weave_b200 import _lib
    from photon_weave_b200.core import ffi

    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    td = torch.complex128 if cd == np.complex128 else torch.complex64
    fill = complex(-1.5, 9.75)
    _lib.set_option("blocks_min_elems", 0)
    _lib.set_option("blocks_min_elems_small_t", 0)
    try:
        for dims, tg, kind in MAT_CASES:
            N = math.prod(dims)
            rho = ops.random_density(N, 4, dtype=cd)
            ibuf, iview = _banded(N * N, td, fill)
            obuf, oview = _banded(N * N, td, fill)
            iview.copy_(torch.as_tensor(rho.reshape(-1)).cuda())
            if kind == "loss":
                K = _operator(kind, dims, tg, cd)
                ffi.kraus_mat(iview, torch.as_tensor(K).cuda(), dims, tg, out=oview)
                ref = oad.apply_kraus_matrix(dims, tg, rho.astype(np.complex128), K.astype(np.complex128), True)
            else:
                op = _operator(kind, dims, tg, cd)
                ffi.apply_mat(iview, torch.as_tensor(op).cuda(), dims, tg, out=oview)
                ref = oad.apply_operation_matrix(dims, tg, rho.astype(np.complex128), op.astype(np.complex128), True)
            torch.cuda.synchronize()
            _check_bands(obuf, N * N, fill, (dims, tg, kind))
            _check_bands(ibuf, N * N, fill, (dims, tg, kind, "input"))
            assert np.array_equal(iview.cpu().numpy(), rho.reshape(-1)), (dims, tg, kind, "input modified")
            assert _rel(oview.cpu().numpy(), ref) < TOL[cd] * 5, (dims, tg, kind)
    finally:
        _lib.set_option("blocks_min_elems", 1 << 18)
        _lib.set_option("blocks_min_elems_small_t", 1 << 22)
