"""
Parity tests proper: the CUDA path (through the C ABI, via the reference-facing
adapters) against the NumPy oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): 1e-12 relative in complex128, 1e-5 in
complex64; measurement outcomes bit-exact under a fixed key.
"""

import math

import numpy as np
import pytest

import seam_cases as sc
from oracle import adapters_np as oad
from oracle import operators as ops
from oracle import rng_threefry as orng

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

TOL = {np.complex128: 1e-12, np.complex64: 1e-5}


@pytest.fixture(scope="module")
def ad():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as g

    g.build()
    from photon_weave_b200.core import adapters

    return adapters


class CudaBackend:
    def __init__(self, ad):
        self.ad = ad

    @staticmethod
    def arr(x):
        return torch.as_tensor(np.asarray(x)).cuda()

    @staticmethod
    def np(x):
        return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)

    @staticmethod
    def key(seed):
        return orng.PRNGKey(seed)


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.abs(b).max()
    return np.abs(a - b).max() / (scale if scale > 0 else 1.0)


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


def host(x):
    return x.detach().cpu().numpy()


# --- the reference's own seam tests, against the CUDA path ---------------------


@pytest.mark.parametrize("use_contraction", [False, True])
@pytest.mark.parametrize("case", sc.CASES_BOTH_ROUTES, ids=lambda c: c.__name__)
def test_seam_case_both_routes(ad, case, use_contraction):
    case(CudaBackend(ad), use_contraction)


@pytest.mark.parametrize("case", sc.CASES_SINGLE, ids=lambda c: c.__name__)
def test_seam_case(ad, case):
    case(CudaBackend(ad))


# --- random-input parity against the oracle ----------------------------------------

SHAPES = [
    ((2, 3, 2), (1,)),
    ((3, 2, 4), (2, 0)),
    ((3, 2, 4), (0, 2)),
    ((2, 3, 2, 2), (3, 1)),
    ((4, 3), (1, 0)),
    ((5,), (0,)),
    ((2, 3, 2, 3), (2, 0, 3)),
    ((6, 6, 6), (1,)),
    ((6, 6, 6), (0, 1)),          # t = 36: generic / tiled path
    ((5, 5, 5), (2, 0)),          # t = 25, reversed order
    ((2, 2, 2, 2, 2, 2, 2), (6,)),
    ((7, 1, 3), (0,)),            # singleton axis
    ((64, 2), (0, 1)),            # Jaynes-Cummings shape: everything is a target, t = 128
    ((3, 4, 5, 2), (1, 2)),
    ((7, 6, 6), (1, 2)),          # innermost contiguous target run (staged block kernel)
    ((4, 3, 6, 6), (3, 2)),       # same, operator order reversed
    ((6, 5, 6), (2, 0)),          # innermost + outermost targets: segmented fibers
    ((3, 6, 2, 6), (3, 1)),
    ((40, 6, 6), (2, 1)),         # more than one warp chunk, partial last chunk
]


@pytest.mark.parametrize("cdtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("dims,targets", SHAPES)
def test_apply_vector_parity(ad, dims, targets, cdtype):
    N = math.prod(dims)
    t = math.prod(dims[i] for i in targets)
    psi = ops.random_state(N, 1, cdtype)
    op = (ops.random_unitary(t, 7) * 1.1).astype(cdtype)
    ref = oad.apply_operation_vector(dims, targets, psi.astype(np.complex128), op.astype(np.complex128), True)
    for uc in (False, True):
        out = host(ad.apply_operation_vector(dims, targets, dev(psi), dev(op), use_contraction=uc))
        assert out.dtype == cdtype and out.shape == (N, 1)
        assert rel_err(out, ref) < TOL[cdtype]


MAT_SHAPES = [
    ((2, 3, 2), (1,)),
    ((3, 2, 4), (2, 0)),
    ((2, 3, 2, 2), (3, 1)),
    ((4, 3), (1, 0)),
    ((5,), (0,)),
    ((5, 5), (0,)),
    ((5, 5), (1, 0)),             # t = 25 on a density matrix
    ((3, 3, 3), (1,)),
    ((8, 8), (1,)),
    ((2, 2, 2, 2), (0, 3)),
]


@pytest.mark.parametrize("cdtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("dims,targets", MAT_SHAPES)
def test_apply_matrix_and_kraus_parity(ad, dims, targets, cdtype):
    N = math.prod(dims)
    t = math.prod(dims[i] for i in targets)
    rho = ops.random_density(N, 2, dtype=cdtype)
    op = (ops.random_unitary(t, 3) * 0.9).astype(cdtype)
    ref = oad.apply_operation_matrix(dims, targets, rho.astype(np.complex128), op.astype(np.complex128), True)
    out = host(ad.apply_operation_matrix(dims, targets, dev(rho), dev(op)))
    assert out.dtype == cdtype and out.shape == (N, N)
    assert rel_err(out, ref) < TOL[cdtype]
    kr = np.stack([ops.random_unitary(t, 10 + k) * (0.3 + 0.2 * k) for k in range(3)]).astype(cdtype)
    refk = oad.apply_kraus_matrix(dims, targets, rho.astype(np.complex128), kr.astype(np.complex128), True)
    for uc in (False, True):
        outk = host(ad.apply_kraus_matrix(dims, targets, dev(rho), dev(kr), use_contraction=uc))
        assert rel_err(outk, refk) < TOL[cdtype]
    # list-of-operators input (state/utils/operations.py:260 stacks a Python list)
    outl = host(ad.apply_kraus_matrix(dims, targets, dev(rho), [dev(k) for k in kr]))
    assert rel_err(outl, refk) < TOL[cdtype]


def test_structured_operators_on_innermost_axes(ad):
    """Beam splitter / shift / diagonal operators (block structured) with the innermost axis
    among the targets, plus a loss channel on the LAST mode of a density matrix."""
    dims = (37, 6, 6)
    N = math.prod(dims)
    psi = ops.random_state(N, 3)
    bs = ops.beam_splitter(6, 6, 0.7)
    shift = np.kron(ops.creation_operator(6), ops.annihilation_operator(6))  # not in-place safe
    diag = np.diag(np.exp(1j * np.arange(36) * 0.1))
    for op, tg in [(bs, (1, 2)), (bs, (2, 1)), (shift, (1, 2)), (diag, (1, 2)), (bs, (2, 0))]:
        t = math.prod(dims[i] for i in tg)
        if op.shape[0] != t:
            op = ops.beam_splitter(6, 37, 0.3) if t == 6 * 37 else op
        out = host(ad.apply_operation_vector(dims, tg, dev(psi), dev(op)))
        assert rel_err(out, oad.apply_operation_vector(dims, tg, psi, op, True)) < 1e-12
    rho = ops.random_density(40 * 8, 4)
    ks = ops.loss_channel(8, 0.15)
    out = host(ad.apply_kraus_matrix((40, 8), (1,), dev(rho), dev(ks)))
    assert rel_err(out, oad.apply_kraus_matrix((40, 8), (1,), rho, ks, True)) < 1e-12
    ph = ops.phase_operator(8, 0.4)
    out = host(ad.apply_operation_matrix((40, 8), (1,), dev(rho), dev(ph)))
    assert rel_err(out, oad.apply_operation_matrix((40, 8), (1,), rho, ph, True)) < 1e-12


def test_loss_channel_preserves_trace_and_matches_oracle(ad):
    dims = (5, 5, 5)
    N = 125
    rho = ops.random_density(N, 4)
    ks = ops.loss_channel(5, 0.1)
    for tg in [(0,), (1,), (2,)]:
        out = host(ad.apply_kraus_matrix(dims, tg, dev(rho), dev(ks)))
        assert abs(np.trace(out) - 1) < 1e-12
        assert rel_err(out, oad.apply_kraus_matrix(dims, tg, rho, ks)) < 1e-12


@pytest.mark.parametrize("cdtype", [np.complex128, np.complex64])
@pytest.mark.parametrize(
    "dims,keep",
    [((2, 3, 2), (0, 2)), ((2, 3, 2), (2, 0)), ((3, 2, 4), (1,)), ((4, 3, 2, 2), (3, 0, 1)),
     ((5, 5), (1,)), ((2, 2, 2), (0, 1, 2)), ((2, 2, 2), (2, 0, 1))],
)
def test_trace_out_parity_both_orders(ad, dims, keep, cdtype):
    N = math.prod(dims)
    rho = ops.random_density(N, 6, dtype=cdtype)
    psi = ops.random_state(N, 8, cdtype)
    for uc in (False, True):
        ref = oad.trace_out_matrix(dims, keep, rho.astype(np.complex128), use_contraction=uc)
        out = host(ad.trace_out_matrix(dims, keep, dev(rho), use_contraction=uc))
        assert rel_err(out, ref) < TOL[cdtype]
        if len(keep) < len(dims):
            refv = oad.trace_out_vector(dims, keep, psi.astype(np.complex128), use_contraction=uc)
            outv = host(ad.trace_out_vector(dims, keep, dev(psi), use_contraction=uc))
            assert rel_err(outv, refv) < TOL[cdtype]


@pytest.mark.parametrize("cdtype", [np.complex128, np.complex64])
@pytest.mark.parametrize(
    "dims,targets",
    [((2, 3, 2), (1,)), ((3, 2, 4), (2, 0)), ((6, 6, 6), (1,)), ((4, 3, 5), (0, 1, 2)),
     ((5, 4), (0,)), ((2, 2, 2, 2, 2), (4, 1))],
)
def test_measure_parity_and_bit_exact_outcomes(ad, dims, targets, cdtype):
    N = math.prod(dims)
    psi = ops.random_state(N, 21, cdtype)
    rho = ops.random_density(N, 22, rank=3, dtype=cdtype)
    tol = TOL[cdtype]
    for seed in range(12):
        key = orng.PRNGKey(seed)
        o_ref, post_ref, probs_ref, _ = oad.measure_vector_with_probs(dims, targets, psi, key)
        o, post, probs, k2 = ad.measure_vector_with_probs(dims, targets, dev(psi), key)
        assert k2 is key
        assert int(o) == int(o_ref)
        assert rel_err(host(probs), probs_ref) < tol
        assert rel_err(host(post), post_ref) < tol * 10
        o_ref, post_ref, probs_ref, _ = oad.measure_matrix_with_probs(dims, targets, rho, key)
        o, post, probs, _ = ad.measure_matrix_with_probs(dims, targets, dev(rho), key)
        assert int(o) == int(o_ref)
        assert rel_err(host(probs), probs_ref) < tol
        assert rel_err(host(post), post_ref) < tol * 10
    # plain variants return the same triple
    o, post, _ = ad.measure_vector(dims, targets, dev(psi), orng.PRNGKey(5))
    o_ref, post_ref, _ = oad.measure_vector(dims, targets, psi, orng.PRNGKey(5))
    assert int(o) == o_ref and rel_err(host(post), post_ref) < tol * 10


def test_reference_seed_kats_on_device(ad):
    """tests/state/test_fock.py:346-380: seed 1 -> 0, seed 3 -> 1 (vector and matrix)."""
    from photon_weave_b200.core import rng

    psi = np.array([[1], [1]], dtype=np.complex128) / np.sqrt(2)
    rho = psi @ psi.conj().T
    for seed, want in ((1, 0), (3, 1)):
        key = rng.split(rng.PRNGKey(seed))[0]
        assert int(ad.measure_vector((2,), (0,), dev(psi), key)[0]) == want
        assert int(ad.measure_matrix((2,), (0,), dev(rho), key)[0]) == want


@pytest.mark.parametrize("cdtype", [np.complex128, np.complex64])
def test_expectation_parity(ad, cdtype):
    dims, targets = (3, 2, 4), (2, 0)
    psi = ops.random_state(24, 31, cdtype)
    rho = ops.random_density(24, 32, dtype=cdtype)
    p_ref, post_ref = oad.measure_vector_expectation(dims, targets, psi.astype(np.complex128))
    p, post = ad.measure_vector_expectation(dims, targets, dev(psi))
    assert rel_err(host(p), p_ref) < TOL[cdtype] and rel_err(host(post), post_ref) < TOL[cdtype]
    p_ref, post_ref = oad.measure_matrix_expectation(dims, targets, rho.astype(np.complex128))
    p, post = ad.measure_matrix_expectation(dims, targets, dev(rho))
    assert rel_err(host(p), p_ref) < TOL[cdtype] and rel_err(host(post), post_ref) < TOL[cdtype]


@pytest.mark.parametrize("dims,targets", [((2, 3, 2), (1,)), ((3, 2, 2), (2, 0)), ((4,), (0,))])
def test_povm_parity(ad, dims, targets):
    N = math.prod(dims)
    t = math.prod(dims[i] for i in targets)
    rho = ops.random_density(N, 41)
    # a 3-outcome POVM from a random unitary's column blocks
    u = ops.random_unitary(t, 42)
    projs = [np.diag((np.arange(t) % 3 == k).astype(float)) for k in range(3)]
    es = np.stack([u @ p @ u.conj().T for p in projs])
    for seed in range(8):
        key = orng.PRNGKey(seed)
        o_ref, post_ref, _ = oad.measure_povm_matrix(dims, targets, es, rho, key)
        o, post, _ = ad.measure_povm_matrix(dims, targets, dev(es), dev(rho), key)
        assert int(o) == int(o_ref)
        assert rel_err(host(post), post_ref) < 1e-11


def test_host_numpy_inputs_round_trip(ad):
    dims, targets = (3, 4, 2), (1,)
    psi = ops.random_state(24, 51)
    op = ops.random_unitary(4, 52)
    out = ad.apply_operation_vector(dims, targets, psi, op)
    assert isinstance(out, np.ndarray)
    assert rel_err(out, oad.apply_operation_vector(dims, targets, psi, op)) < 1e-12
    o, post, _ = ad.measure_vector(dims, targets, psi, orng.PRNGKey(2))
    assert isinstance(o, int) and isinstance(post, np.ndarray)


def test_circuit_host_entry_point(ad):
    from photon_weave_b200.core import ffi

    dims = (4, 3, 5)
    psi = ops.random_state(60, 61)
    o1, o2 = ops.random_unitary(3, 62), ops.random_unitary(20, 63)
    out = ffi.circuit_host(psi, False, dims, [(0, (1,), o1), (0, (2, 0), o2)])
    ref = oad.apply_operation_vector(dims, (2, 0), oad.apply_operation_vector(dims, (1,), psi, o1), o2)
    assert rel_err(out, ref) < 1e-12
    rho = ops.random_density(12, 64)
    ks = ops.loss_channel(3, 0.2)
    u = ops.random_unitary(4, 65)
    outm = ffi.circuit_host(rho, True, (4, 3), [(1, (0,), u), (2, (1,), ks)])
    refm = oad.apply_kraus_matrix((4, 3), (1,), oad.apply_operation_matrix((4, 3), (0,), rho, u), ks)
    assert rel_err(outm, refm) < 1e-12


# --- size-independent properties at larger sizes -------------------------------------


def test_unitary_preserves_norm_and_inverse_round_trip_large(ad):
    from photon_weave_b200.core import ffi

    dims = (6,) * 8  # 1.68M amplitudes
    N = 6**8
    g = torch.Generator(device="cuda").manual_seed(0)
    psi = torch.randn(N, 2, generator=g, device="cuda", dtype=torch.float64)
    psi = torch.view_as_complex(psi).reshape(N, 1)
    n0 = float(ffi.norm_stats(psi)[0])
    cur = psi
    us = []
    for ax in range(8):
        u = ops.random_unitary(6, 100 + ax)
        us.append(u)
        cur = ad.apply_operation_vector(dims, (ax,), cur, dev(u))
    bs = ops.beam_splitter(6, 6, np.pi / 4)
    cur = ad.apply_operation_vector(dims, (2, 5), cur, dev(bs))
    assert abs(float(ffi.norm_stats(cur)[0]) - n0) / n0 < 1e-12
    cur = ad.apply_operation_vector(dims, (2, 5), cur, dev(bs.conj().T))
    for ax in reversed(range(8)):
        cur = ad.apply_operation_vector(dims, (ax,), cur, dev(us[ax].conj().T))
    assert float((cur - psi).abs().max()) < 1e-12 * float(psi.abs().max()) * 50


def test_linearity_and_sample_check_large(ad):
    dims = (6,) * 7
    N = 6**7
    a = ops.random_state(N, 71)
    b = ops.random_state(N, 72)
    op = ops.displacement_operator(6, 0.5)
    fa = host(ad.apply_operation_vector(dims, (3,), dev(a), dev(op)))
    fb = host(ad.apply_operation_vector(dims, (3,), dev(b), dev(op)))
    fab = host(ad.apply_operation_vector(dims, (3,), dev(2 * a - 1j * b), dev(op)))
    assert rel_err(fab, 2 * fa - 1j * fb) < 1e-12
    assert rel_err(fa, oad.apply_operation_vector(dims, (3,), a, op)) < 1e-12


@pytest.mark.parametrize("cdtype", [np.complex128, np.complex64])
@pytest.mark.parametrize(
    "dims,perm",
    [((2, 3, 4), (2, 0, 1)), ((3, 2, 2, 5), (1, 3, 0, 2)), ((6, 6, 6), (0, 2, 1)), ((4, 1, 3), (2, 1, 0)),
     ((2, 3, 2, 3), (0, 1, 2, 3))],
)
def test_permute_matches_numpy_transpose(ad, dims, perm, cdtype):
    """pw_permute == jnp.transpose of reorder_vector_front (core/adapters.py:228-230)."""
    from photon_weave_b200.core import ffi

    N = math.prod(dims)
    x = ops.random_state(N, 3, cdtype).reshape(-1)
    out = host(ffi.permute(dev(x), dims, perm))
    ref = np.ascontiguousarray(np.transpose(x.reshape(dims), perm)).reshape(-1)
    assert np.array_equal(out, ref)


def test_sharded_state_single_rank_cuda_backend(ad):
    """World size 1 degenerates to the plain kernels (exercises CudaBackend wiring)."""
    from photon_weave_b200.sharded import CudaBackend, ShardedStateVector, VAxis

    dims = (6, 6, 6)
    psi = ops.random_state(216, 9)
    vaxes = [VAxis(g, d, 1) for g, d in enumerate(dims)]
    st = ShardedStateVector(dims, dev(psi).reshape(-1), vaxes, [], 0, 1, None, CudaBackend())
    bs = ops.beam_splitter(6, 6, np.pi / 4)
    st.apply_operation(dev(bs), (2, 0))
    ref = oad.apply_operation_vector(dims, (2, 0), psi, bs)
    assert rel_err(host(st.local).reshape(-1, 1), ref) < 1e-12


def test_gather_kernels_single_process_emulation(ad):
    """pw_gather_chunks / pw_apply_vec_gather with W chunks living on ONE GPU (the peer
    pointers are ordinary device pointers here): fused gather+apply == apply(concat)."""
    from photon_weave_b200.sharded import CudaBackend

    be = CudaBackend()
    W, dims = 4, (4, 3, 6, 5)          # virtual tensor: leading axis of extent W
    N = math.prod(dims)
    chunk = N // W
    psi = ops.random_state(N, 77).reshape(-1)
    pieces = [dev(psi[p * chunk:(p + 1) * chunk].copy()) for p in range(W)]
    srcs = [t.data_ptr() for t in pieces]
    out = torch.empty(N, dtype=torch.complex128, device="cuda")
    be.gather_chunks(srcs, chunk, out, rank=2)
    assert np.array_equal(host(out), psi)
    for targets in [(2,), (0,), (0, 1), (3, 0)]:
        t = math.prod(dims[i] for i in targets)
        if t > 8:
            continue
        op = ops.random_unitary(t, 5)
        ok = be.apply_vec_gather(srcs, chunk, out, dev(op), dims, targets, rank=3)
        assert ok
        ref = oad.apply_operation_vector(dims, targets, psi.reshape(-1, 1), op)
        assert rel_err(host(out).reshape(-1, 1), ref) < 1e-12
    # dense t > 8: not covered (gather, then apply); block-structured t > 8: fused
    assert be.apply_vec_gather(srcs, chunk, out, dev(ops.random_unitary(12, 1)), dims, (0, 1)) is False
    for targets, op in [((2, 3), ops.beam_splitter(6, 5, 0.7)), ((3, 2), ops.beam_splitter(5, 6, 0.3)),
                        ((0, 2), np.kron(ops.phase_operator(4, 0.3), ops.phase_operator(6, 1.1))),
                        ((0, 3), ops.beam_splitter(4, 5, 0.9))]:
        assert be.apply_vec_gather(srcs, chunk, out, dev(op), dims, targets, rank=1)
        ref = oad.apply_operation_vector(dims, targets, psi.reshape(-1, 1), op)
        assert rel_err(host(out).reshape(-1, 1), ref) < 1e-12, targets


@pytest.mark.parametrize("cdtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("dims,target", [((6, 6), (1,)), ((5, 4, 3, 8), (3,)), ((3, 2, 4, 2), (2, 3)),
                                         ((3, 2, 2, 4), (3, 2)), ((7, 5, 2), (2,))])
def test_contiguous_fiber_kernel_parity(ad, dims, target, cdtype):
    """Targets on the innermost axes: warp-transposing register kernel (needs >= 4096 fibers
    to be selected, so the leading axis is inflated)."""
    big = (4096,) + tuple(dims)
    tg = tuple(t + 1 for t in target)
    N = math.prod(big)
    t = math.prod(big[i] for i in tg)
    psi = ops.random_state(N, 5, cdtype)
    op = (ops.random_unitary(t, 6) * 0.9).astype(cdtype)
    ref = oad.apply_operation_vector(big, tg, psi.astype(np.complex128), op.astype(np.complex128), True)
    out = host(ad.apply_operation_vector(big, tg, dev(psi), dev(op)))
    assert rel_err(out, ref) < TOL[cdtype]
    rho_dims = (64,) + tuple(dims[-2:])
    D = math.prod(rho_dims)
    rho = ops.random_density(D, 7, dtype=cdtype)
    last = len(rho_dims) - 1
    d = rho_dims[last]
    u = (ops.random_unitary(d, 8)).astype(cdtype)
    refm = oad.apply_operation_matrix(rho_dims, (last,), rho.astype(np.complex128), u.astype(np.complex128), True)
    outm = host(ad.apply_operation_matrix(rho_dims, (last,), dev(rho), dev(u)))
    assert rel_err(outm, refm) < TOL[cdtype]


@pytest.mark.parametrize("cdtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("d", [2, 3, 4, 5, 6])
def test_fused_pair_and_device_circuit(ad, d, cdtype):
    from photon_weave_b200.core import ffi

    dims = (d, 3, d, 2, d)
    N = math.prod(dims)
    psi = ops.random_state(N, 9, cdtype)
    ua, ub = ops.random_unitary(d, 1).astype(cdtype), ops.random_unitary(d, 2).astype(cdtype)
    for (a, b) in [(0, 2), (4, 0), (2, 4)]:
        out = ffi.apply_vec_pair(dev(psi), dev(ua), dev(ub), dims, a, b)
        assert out is not None
        ref = oad.apply_operation_vector(dims, (b,), oad.apply_operation_vector(
            dims, (a,), psi.astype(np.complex128), ua.astype(np.complex128)), ub.astype(np.complex128))
        assert rel_err(host(out).reshape(-1, 1), ref) < TOL[cdtype]
    if d != 3:  # axes of different dimension are not fusable
        assert ffi.apply_vec_pair(dev(psi), dev(ua), dev(ops.random_unitary(3, 1).astype(cdtype)), dims, 0, 1) is None
    # device-resident circuit, fused and unfused, equals op-by-op application
    u3 = ops.random_unitary(3, 3).astype(cdtype)
    bs = ops.beam_splitter(d, d, 0.4).astype(cdtype)
    prog = [(0, (0,), ua), (0, (2,), ub), (0, (1,), u3), (0, (4,), ua), (0, (2, 4), bs), (0, (2,), ub)]
    ref = psi.astype(np.complex128)
    for _, tg, o in prog:
        ref = oad.apply_operation_vector(dims, tg, ref, o.astype(np.complex128), True)
    for fuse in (False, True):
        st, sc = dev(psi).reshape(-1).clone(), torch.empty(N, dtype=dev(psi).dtype, device="cuda")
        res, _ = ffi.circuit_device(st, sc, False, dims, [(k, tg, dev(o)) for k, tg, o in prog], fuse=fuse)
        assert rel_err(host(res).reshape(-1, 1), ref) < TOL[cdtype] * 5


@pytest.mark.parametrize("cdtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("dims,keep", [((64, 2), (1,)), ((3, 5, 2), (2, 0)), ((7,), (0,)), ((16, 8), (1,)), ((33,), (0,))][:4])
def test_fused_evolve_matches_stepwise(ad, dims, keep, cdtype):
    """pw_evolve_vec == nsteps x (apply_operation on all subsystems, trace_out of `keep`)."""
    from photon_weave_b200.core import ffi

    N = math.prod(dims)
    U = ops.random_unitary(N, 3).astype(cdtype)
    psi = ops.random_state(N, 4, cdtype)
    nsteps = 25
    out, red = ffi.evolve_vec(dev(psi).reshape(-1), dev(U), dims, nsteps, keep=keep)
    ref = psi.astype(np.complex128)
    tol = TOL[cdtype] * (20 if cdtype == np.complex64 else 50)
    for s in range(nsteps):
        ref = U.astype(np.complex128) @ ref
        want = oad.trace_out_vector(dims, keep, ref, use_contraction=False)
        assert rel_err(host(red[s]), want) < tol
    assert rel_err(host(out).reshape(-1, 1), ref) < tol
    # without observables
    out2, red2 = ffi.evolve_vec(dev(psi).reshape(-1), dev(U), dims, nsteps)
    assert red2 is None and rel_err(host(out2).reshape(-1, 1), ref) < tol


def test_small_state_dense_rows_path(ad):
    """Launch-latency-bound shape (Jaynes-Cummings, t = 128, one fiber) and a t = 25 case."""
    for dims, tg in [((64, 2), (0, 1)), ((5, 5, 3), (1, 0)), ((2, 36), (1,))]:
        N = math.prod(dims)
        t = math.prod(dims[i] for i in tg)
        psi = ops.random_state(N, 1)
        op = ops.random_unitary(t, 2)
        out = host(ad.apply_operation_vector(dims, tg, dev(psi), dev(op)))
        assert rel_err(out, oad.apply_operation_vector(dims, tg, psi, op)) < 1e-12


def test_layout_ops_parity(ad):
    """pw_resize_axis / pw_kron / pw_expand_vec / pw_contract_mat against the oracle restatement
    of resize_fock, kron_reduce, state_expand and state_contract."""
    from oracle import layout_np as lay
    from photon_weave_b200.core import ffi

    dims = (3, 4, 2)
    psi = ops.random_state(24, 1)
    rho = ops.random_density(24, 2)
    for axis, nd in [(1, 6), (1, 2), (0, 5), (2, 1), (2, 4)]:
        out = host(ffi.resize_axis(dev(psi), dims, axis, nd, False))
        assert np.array_equal(out, lay.resize_fock_vector(dims, axis, nd, psi))
        outm = host(ffi.resize_axis(dev(rho), dims, axis, nd, True))
        assert np.array_equal(outm, lay.resize_fock_matrix(dims, axis, nd, rho))
    a, b = ops.random_state(5, 3), ops.random_state(7, 4)
    assert rel_err(host(ffi.kron(dev(a), dev(b))), np.kron(a, b)) < 1e-15
    A_, B_ = ops.random_density(4, 5), ops.random_density(3, 6)
    assert rel_err(host(ffi.kron(dev(A_), dev(B_))), np.kron(A_, B_)) < 1e-15
    assert rel_err(host(ffi.expand_vec(dev(psi))), lay.state_expand_vector(psi)) < 1e-15
    # contraction of a pure state: same vector and phase convention as eigh + phase removal
    pure = lay.state_expand_vector(psi)
    got, info, flag = ffi.contract_mat(dev(pure))
    want, ok = lay.state_contract_matrix(pure)
    assert ok and int(flag) == 1
    assert abs(float(info[0]) - 1) < 1e-12 and abs(float(info[1]) - 1) < 1e-12
    assert rel_err(host(got), want) < 1e-12
    got, info, flag = ffi.contract_mat(dev(rho))
    assert int(flag) == 0 and not lay.state_contract_matrix(rho)[1]
    assert abs(float(info[0]) - np.trace(rho @ rho).real) < 1e-12


@pytest.mark.parametrize(
    "dims",
    [(4, 3, 5), (64, 2), (6,) * 6, (3, 40000), (20000,), (2, 9000, 3), (5, 5, 5, 5), (7, 6, 5, 4, 3, 2, 2, 3)],
)
def test_occupancy_matches_reference_loop(dims):
    """pw_occupancy == trace_out -> num_quanta per subsystem (composite_envelope.py:688-696)."""
    from oracle import layout_np as lay
    from photon_weave_b200.core import ffi

    rng = np.random.default_rng(len(dims))
    N = int(np.prod(dims))
    for trial in range(3):
        # a random sub-box of occupied levels, sparse inside it
        hi = [int(rng.integers(0, d)) for d in dims]
        t = np.zeros(dims, dtype=np.complex128)
        box = tuple(slice(0, h + 1) for h in hi)
        vals = rng.standard_normal(t[box].shape) + 1j * rng.standard_normal(t[box].shape)
        vals *= rng.random(t[box].shape) < (0.3 if trial else 1.0)
        t[box] = vals
        t[tuple(hi)] = 0.5 - 0.25j  # make the corner occupied so the expected answer is hi
        psi = t.reshape(-1, 1)
        levels, stats = ffi.occupancy(dev(psi), dims, False)
        assert levels.tolist() == hi
        assert abs(float(stats[0]) - np.vdot(psi, psi).real) < 1e-9 * max(1.0, np.vdot(psi, psi).real)
        assert abs(float(stats[1]) - np.max(np.abs(psi) ** 2)) < 1e-12 * np.max(np.abs(psi) ** 2)
        if N <= 4096:
            assert levels.tolist() == lay.occupancy_levels(dims, psi, False)
            A_ = rng.standard_normal((N, 2)) + 1j * rng.standard_normal((N, 2))
            A_ *= (np.abs(psi) > 0)
            rho = A_ @ A_.conj().T
            lm, sm = ffi.occupancy(dev(rho), dims, True)
            assert lm.tolist() == lay.occupancy_levels(dims, rho, True)
            assert abs(float(sm[0]) - np.trace(rho).real) < 1e-9 * np.trace(rho).real
    z = np.zeros((N, 1), dtype=np.complex64)
    levels, stats = ffi.occupancy(dev(z), dims, False)
    assert levels.tolist() == [-1] * len(dims) and float(stats[0]) == 0.0


def test_host_pipeline_matches_serial_path():
    """pw_pipeline_submit/wait: several jobs in flight with different inputs and operator
    lists give exactly what pw_circuit_host gives, job by job."""
    import torch

    from photon_weave_b200.core import ffi

    dims = (3, 4, 5, 2)
    n = int(np.prod(dims))
    jobs = []
    for j in range(7):
        psi = ops.random_state(n, 100 + j)
        oplist = [(0, (j % 4,), ops.random_unitary(dims[j % 4], j)),
               (0, (3, 1), ops.random_unitary(8, 50 + j))]
        if j % 2:
            oplist.append((0, (2,), ops.random_unitary(5, 70 + j)))  # odd count: result in the scratch
        jobs.append((psi, oplist))
    want = [ffi.circuit_host(psi, False, dims, oplist) for psi, oplist in jobs]
    for nbuf in (2, 3, 4):
        ins = [torch.as_tensor(psi.reshape(-1)).pin_memory().numpy() for psi, _ in jobs]
        outs = [torch.empty(n, dtype=torch.complex128).pin_memory().numpy() for _ in jobs]
        with ffi.HostPipeline(dims, False, np.complex128, buffers=nbuf) as pl:
            assert pl.buffers == nbuf
            tickets = [pl.submit(ins[j], outs[j], jobs[j][1]) for j in range(len(jobs))]
            for tk in reversed(tickets):
                pl.wait(tk)
        for j in range(len(jobs)):
            assert np.array_equal(outs[j].reshape(-1), want[j].reshape(-1)), (nbuf, j)
    rho = ops.random_density(12, 5)
    ks = ops.loss_channel(3, 0.2)
    wantm = ffi.circuit_host(rho, True, (4, 3), [(1, (0,), ops.random_unitary(4, 1)), (2, (1,), ks)])
    outm = np.empty_like(rho)
    with ffi.HostPipeline((4, 3), True, np.complex128, buffers=2) as pl:
        pl.wait(pl.submit(np.ascontiguousarray(rho), outm,
                          [(1, (0,), ops.random_unitary(4, 1)), (2, (1,), ks)]))
    assert np.array_equal(outm, wantm)


@pytest.mark.parametrize("cdtype", [np.complex128, np.complex64])
def test_two_sided_single_pass_kernel(ad, cdtype):
    """U rho U^dagger through the shared-memory two-sided kernel (dense d <= 8 operators and
    block-structured two-mode operators, innermost axis untouched), including ragged fiber
    counts, operator-order targets and the operators it must hand to the fallbacks (zero rows,
    diagonal, blocks larger than 8)."""
    from photon_weave_b200 import _lib

    tol = TOL[cdtype]

    def check(dims, targets, op, seed):
        N = math.prod(dims)
        rho = ops.random_density(N, seed, dtype=cdtype)
        ref = oad.apply_operation_matrix(dims, targets, rho.astype(np.complex128),
                                         np.asarray(op, dtype=np.complex128), True)
        out = host(ad.apply_operation_matrix(dims, targets, dev(rho), dev(np.asarray(op, dtype=cdtype))))
        assert rel_err(out, ref) < tol, (dims, targets)

    # (thresholds off: every shape below takes the structured / tensor-core paths)
    _lib.set_option("blocks_min_elems", 0)
    try:
        # large enough for the structured paths without touching the tuning knobs
        check((8, 8, 10), (0,), ops.random_unitary(8, 1), 1)
        check((8, 8, 10), (1,), ops.displacement_operator(8, 0.4), 2)
        check((8, 8, 10), (0, 1), ops.beam_splitter(8, 8, 0.6), 3)
        check((8, 8, 10), (1, 0), ops.beam_splitter(8, 8, 0.6), 4)
        check((6, 5, 7, 3), (1,), ops.random_unitary(5, 5), 5)
        check((6, 5, 7, 3), (2, 0), ops.beam_splitter(7, 6, 0.3), 6)
        # innermost axis IS the target (tensor-core kernel with the transposing load, complex128;
        # complex64 and ragged runs take the two-pass route)
        check((8, 8, 8), (2,), ops.random_unitary(8, 21), 21)
        check((4, 16, 6), (2,), ops.random_unitary(6, 22), 22)
        check((2, 32, 5), (2,), ops.squeezing_operator(5, 0.2 + 0.2j), 23)
        check((64, 7), (1,), ops.random_unitary(7, 24), 24)
        check((3, 24, 8), (2,), ops.random_unitary(8, 25), 25)           # run of 72: ragged wide tiles
        check((9, 11, 6), (2,), ops.random_unitary(6, 28), 28)
        # two-mode operators on the innermost modes (wide-tile tensor-core kernel)
        check((3, 8, 8), (1, 2), ops.beam_splitter(8, 8, 0.4), 29)
        check((3, 8, 8), (2, 1), ops.beam_splitter(8, 8, 0.4), 30)
        check((5, 5, 5, 5), (2, 3), ops.beam_splitter(5, 5, 1.1), 31)
        check((4, 3, 7, 6), (3, 2), ops.beam_splitter(6, 7, 0.8), 32)
        check((10, 4, 7, 2, 4), (2, 3, 4), np.kron(ops.random_unitary(7, 33), ops.random_unitary(8, 34)), 33)
        # innermost + far column targets: segmented fiber rows
        check((8, 3, 8), (2, 0), ops.beam_splitter(8, 8, 0.5), 35)
        check((8, 3, 8), (0, 2), ops.beam_splitter(8, 8, 0.5), 36)
        check((4, 9, 2, 6), (3, 0), ops.beam_splitter(6, 4, 0.6), 37)
        check((5, 2, 7, 5), (0, 3), ops.beam_splitter(5, 5, 0.2), 38)
        check((3, 2, 5, 4, 3), (4, 0, 3), np.kron(np.kron(ops.random_unitary(3, 39), np.eye(3)), ops.random_unitary(4, 40)), 39)
        # beam splitters whose sectors pack into several super-blocks (cutoffs 8, 5, 7 x 6)
        check((3, 8, 8, 4), (1, 2), ops.beam_splitter(8, 8, 0.9), 26)
        check((5, 5, 5, 5), (3, 1), ops.beam_splitter(5, 5, 0.7), 27)
        for tg in [(0,), (1,), (0, 1), (1, 0)]:
            t = math.prod((3, 4, 5)[i] for i in tg)
            check((3, 4, 5), tg, ops.random_unitary(t, 7) if t <= 8 else
                  ops.beam_splitter(*[(3, 4, 5)[i] for i in tg], 0.5), 7)
        check((4, 6, 9), (1,), ops.annihilation_operator(6) + 0.1 * np.eye(6), 8)
        check((4, 6, 9), (1,), ops.annihilation_operator(6), 9)          # zero row: fallback
        check((4, 6, 9), (1,), ops.phase_operator(6, 0.3), 10)           # diagonal: scalar blocks
        check((4, 4, 9), (0, 1), ops.random_unitary(16, 11), 11)         # one 16 x 16 block: fallback
        check((2, 2, 2, 7), (2, 0), ops.random_unitary(4, 12), 12)
    finally:
        _lib.set_option("blocks_min_elems", 1 << 18)
    # default thresholds: a single t <= 8 operator on a small rho skips the analysis (two
    # register-kernel passes), a beam splitter of the same size takes the structured path
    check((8, 8, 10), (0,), ops.random_unitary(8, 1), 1)
    check((8, 8, 10), (0, 1), ops.beam_splitter(8, 8, 0.6), 3)


def test_dense_tensor_core_kernel(ad):
    """Dense complex128 operators with 8 < t <= 64 on adjacent fibers take the FP64 tensor-core
    kernel (pw_dense.cu): vectors, the conjugated column side of a density matrix, ragged fiber
    counts, operator-order targets, t not a multiple of 4 or 8."""
    from photon_weave_b200 import _lib

    tol = TOL[np.complex128]

    def check_vec(dims, targets, op, seed):
        N = math.prod(dims)
        psi = ops.random_state(N, seed)
        ref = oad.apply_operation_vector(dims, targets, psi, np.asarray(op, dtype=np.complex128), True)
        out = host(ad.apply_operation_vector(dims, targets, dev(psi), dev(np.asarray(op, dtype=np.complex128))))
        assert rel_err(out, ref) < tol, (dims, targets)

    def check_mat(dims, targets, op, seed):
        N = math.prod(dims)
        rho = ops.random_density(N, seed)
        ref = oad.apply_operation_matrix(dims, targets, rho, np.asarray(op, dtype=np.complex128), True)
        out = host(ad.apply_operation_matrix(dims, targets, dev(rho), dev(np.asarray(op, dtype=np.complex128))))
        assert rel_err(out, ref) < tol, (dims, targets)

    _lib.set_option("blocks_min_elems", 0)
    try:
        check_vec((6, 6, 50), (0, 1), ops.random_unitary(36, 41), 41)
        check_vec((6, 6, 50), (1, 0), ops.random_unitary(36, 42), 42)
        check_vec((3, 8, 8, 77), (1, 2), ops.random_unitary(64, 43), 43)
        check_vec((9, 3, 5, 33), (0,), ops.random_unitary(9, 44), 44)
        check_vec((2, 13, 2, 19), (1,), ops.displacement_operator(13, 0.3 + 0.1j), 45)
        check_vec((5, 7, 16), (1, 0), ops.random_unitary(35, 46), 46)
        check_vec((3, 3, 3, 40), (2, 0, 1), ops.random_unitary(27, 47), 47)
        check_mat((4, 4, 20), (0, 1), ops.random_unitary(16, 48), 48)
        # targets = the innermost axes (contiguous fibers): transposing tile layout
        check_vec((50, 6, 6), (1, 2), ops.random_unitary(36, 51), 51)
        check_vec((50, 6, 6), (2, 1), ops.random_unitary(36, 52), 52)
        check_vec((7, 13), (1,), ops.random_unitary(13, 53), 53)
        check_vec((3, 70, 8, 8), (2, 3), ops.random_unitary(64, 54), 54)
        # Kraus sum with a dense superoperator (8^2 = 64): one dense pass on the doubled tensor
        for dims_k, tg_k, dk in [((8, 70), (0,), 8), ((3, 6, 40), (1,), 6), ((5, 2, 2, 33), (1, 2), 4)]:
            rng = np.random.default_rng(57)
            ks = rng.standard_normal((3, dk, dk)) + 1j * rng.standard_normal((3, dk, dk))
            ks /= np.sqrt(np.trace(sum(k.conj().T @ k for k in ks)).real / dk)
            Nk = math.prod(dims_k)
            rho = ops.random_density(Nk, 58)
            ref = oad.apply_kraus_matrix(dims_k, tg_k, rho, ks, True)
            out = host(ad.apply_kraus_matrix(dims_k, tg_k, dev(rho), dev(ks)))
            assert rel_err(out, ref) < tol, (dims_k, tg_k)
        check_mat((20, 4, 4), (1, 2), ops.random_unitary(16, 55), 55)
        check_mat((33, 11), (1,), ops.random_unitary(11, 56), 56)
        check_mat((3, 5, 16), (1, 0), ops.random_unitary(15, 49), 49)
    finally:
        _lib.set_option("blocks_min_elems", 1 << 18)


def test_staged_block_kernel_on_tensor_cores(ad):
    """Block-structured complex128 operators whose innermost axis is a target take the staged
    kernel's tensor-core variant (super-blocks of <= 8 indices, pw_blocks.cu): beam splitters on
    the innermost modes, innermost + far targets (segmented fibers), loss channels on the
    innermost mode of a density matrix (superoperator blocks), ragged fiber counts."""
    from photon_weave_b200 import _lib

    tol = TOL[np.complex128]
    _lib.set_option("blocks_min_elems", 0)
    try:
        for seed, (dims, targets, op) in enumerate([
            ((70, 6, 6), (1, 2), ops.beam_splitter(6, 6, 0.7)),
            ((70, 6, 6), (2, 1), ops.beam_splitter(6, 6, 0.7)),
            ((6, 100, 6), (2, 0), ops.beam_splitter(6, 6, 1.2)),
            ((6, 100, 6), (0, 2), ops.beam_splitter(6, 6, 1.2)),
            ((8, 3, 33, 8), (0, 3), ops.beam_splitter(8, 8, 0.3)),
            ((5, 77, 7), (2, 0), ops.beam_splitter(7, 5, 0.9)),
            ((3, 4, 5, 4), (3, 1), np.kron(ops.random_unitary(4, 1), np.eye(4))[:, np.random.default_rng(3).permutation(16)]),
        ]):
            N = math.prod(dims)
            psi = ops.random_state(N, 60 + seed)
            ref = oad.apply_operation_vector(dims, targets, psi, np.asarray(op, dtype=np.complex128), True)
            out = host(ad.apply_operation_vector(dims, targets, dev(psi), dev(np.asarray(op, dtype=np.complex128))))
            assert rel_err(out, ref) < tol, (dims, targets)
        for seed, (dims, targets, d) in enumerate([((70, 8), (1,), 8), ((5, 9, 6), (2,), 6), ((130, 5), (1,), 5)]):
            N = math.prod(dims)
            rho = ops.random_density(N, 70 + seed)
            ks = ops.loss_channel(d, 0.2)
            ref = oad.apply_kraus_matrix(dims, targets, rho, ks, True)
            out = host(ad.apply_kraus_matrix(dims, targets, dev(rho), dev(ks)))
            assert rel_err(out, ref) < tol, (dims, targets)
    finally:
        _lib.set_option("blocks_min_elems", 1 << 18)


def test_contiguous_fibers_on_tensor_cores(ad):
    """One 5..8-level complex128 operator on the innermost axis (contiguous fibers): the
    direct-from-global tensor-core kernel (transposed product, 256-bit stores), including odd t
    (unaligned pairs), ragged fiber counts and the conjugated column pass of a density matrix."""
    tol = TOL[np.complex128]
    for seed, (dims, d) in enumerate([((5000, 6), 6), ((4101, 5), 5), ((4099, 7), 7), ((70, 61, 8), 8),
                                      ((3, 1500, 6), 6)]):
        N = math.prod(dims)
        psi = ops.random_state(N, 80 + seed)
        op = ops.random_unitary(d, 90 + seed) * (1.0 + 0.1 * seed)
        tg = (len(dims) - 1,)
        ref = oad.apply_operation_vector(dims, tg, psi, op, True)
        out = host(ad.apply_operation_vector(dims, tg, dev(psi), dev(op)))
        assert rel_err(out, ref) < tol, dims
    for seed, (dims, d) in enumerate([((70, 8), 8), ((90, 6), 6), ((75, 7), 7)]):
        N = math.prod(dims)
        rho = ops.random_density(N, 95 + seed)
        op = ops.random_unitary(d, 97 + seed)
        ref = oad.apply_operation_matrix(dims, (1,), rho, op, True)
        out = host(ad.apply_operation_matrix(dims, (1,), dev(rho), dev(op)))
        assert rel_err(out, ref) < tol, dims


@pytest.mark.parametrize("cdtype", [np.complex128, np.complex64])
def test_diagonal_operator_streaming_kernel(ad, cdtype):
    """Diagonal operators on large vectors (>= 2^22 amplitudes) take the streaming kernel
    (device-side test of the operator, element pairs, several targets); a non-diagonal
    operator of the same shape must fall through to the register kernel."""
    tol = TOL[cdtype]
    dims = (6,) * 8 + (4,)                      # 6.7 M amplitudes; targets with a short inner run take the kernel
    N = math.prod(dims)
    psi = ops.random_state(N, 3, cdtype)
    ph6, ph4 = ops.phase_operator(6, 0.37), ops.phase_operator(4, -1.1)
    cases = [((3,), ph6), ((0,), ph6), ((7,), ph6), ((6,), ph6), ((5,), ph6), ((8,), ph4), ((2, 7), np.kron(ph6, ph6)),
             ((8, 1), np.diag(np.exp(1j * np.arange(24) * 0.1))),     # t = 24 > 8: not the streaming kernel
             ((4,), ops.displacement_operator(6, 0.3))]
    for tg, op in cases:
        t = math.prod(dims[i] for i in tg)
        op = np.asarray(op, dtype=np.complex128)
        assert op.shape == (t, t)
        ref = oad.apply_operation_vector(dims, tg, psi.astype(np.complex128), op, True)
        out = host(ad.apply_operation_vector(dims, tg, dev(psi), dev(op.astype(cdtype))))
        assert rel_err(out, ref) < tol, tg


@pytest.mark.parametrize("cdtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("dims,axes", [
    ((6, 6, 6, 6, 6, 6), (0, 1, 2)), ((6, 6, 6, 6, 6, 6), (3, 4, 5)), ((6, 6, 6, 6, 6, 6), (4, 1, 5)),
    ((6, 6, 6, 6, 6, 6), (2, 0)), ((6, 6, 6, 6, 6), (0, 1, 2, 3)), ((6, 6, 6, 6, 6), (1, 2, 3, 4)),
    ((4, 3, 4, 5, 4, 8), (2, 0, 4)), ((2,) * 10, (9, 3, 0, 5)), ((8, 8, 8, 4), (1, 0, 2)), ((3, 3, 3), (0, 1, 2)),
    ((5, 7, 5, 16), (0, 2)), ((6, 6, 36), (1, 0)),
])
def test_fused_multi_operator_tile_kernel(ad, dims, axes, cdtype):
    """pw_apply_vec_multi: k single-subsystem operators on different axes in ONE pass equal their
    one-by-one application (core/kernels.py:62-98 applied k times); in place too."""
    from photon_weave_b200.core import ffi

    N = math.prod(dims)
    psi = ops.random_state(N, 61, cdtype)
    us = [ops.random_unitary(dims[a], 70 + i).astype(cdtype) for i, a in enumerate(axes)]
    ref = psi.astype(np.complex128)
    for a, u in zip(axes, us):
        ref = oad.apply_operation_vector(dims, (a,), ref, u.astype(np.complex128))
    d_psi = dev(psi).reshape(-1)
    out = ffi.apply_vec_multi(d_psi, [dev(u) for u in us], dims, axes)
    # the tile kernel declines shapes whose tile rows would not be a contiguous piece of >= 64 bytes
    # (>= 32 when the innermost axis itself is a target); the caller then applies them one by one
    es = 16 if cdtype == np.complex128 else 8
    d, k, n = dims[axes[0]], len(axes), len(dims)
    cmax = (28 * 1024) // (es * d**k)
    if (n - 1) in axes:
        supported = cmax >= 1 and d * es >= 32
    else:
        run = math.prod(dims[max(axes) + 1:])
        c = max([c for c in range(1, cmax + 1) if run % c == 0], default=0)
        supported = c * es >= 64
    if not supported:
        assert out is None
        return
    assert out is not None, "shape should be supported"
    assert rel_err(host(out).reshape(-1, 1), ref) < TOL[cdtype]
    same = ffi.apply_vec_multi(d_psi, [dev(u) for u in us], dims, axes, out=d_psi)
    assert rel_err(host(same).reshape(-1, 1), ref) < TOL[cdtype]


def test_fused_multi_operator_rejects_what_it_cannot_tile(ad):
    from photon_weave_b200.core import ffi

    psi = dev(ops.random_state(4 * 3 * 4, 1)).reshape(-1)
    u4, u3 = dev(ops.random_unitary(4, 1)), dev(ops.random_unitary(3, 2))
    assert ffi.apply_vec_multi(psi, [u4, u3], (4, 3, 4), (0, 1)) is None          # mixed dimensions
    with pytest.raises(Exception):
        ffi.apply_vec_multi(psi, [u4, u4], (4, 3, 4), (0, 0))                       # duplicate axis


def test_device_circuit_groups_commuting_operators(ad):
    """pw_circuit_device with fuse: a layer of single-mode operators + beam splitters equals the
    op-by-op result (the program order is kept wherever operators share a subsystem)."""
    from photon_weave_b200 import _lib
    from photon_weave_b200.core import ffi

    dims = (6,) * 7
    N = math.prod(dims)
    psi = ops.random_state(N, 5)
    layer = [((m,), ops.random_unitary(6, 10 + m)) for m in range(7)]
    layer += [((0, 1), ops.beam_splitter(6, 6, 0.7)), ((3,), ops.random_unitary(6, 30)), ((4,), ops.random_unitary(6, 31)),
              ((3,), ops.random_unitary(6, 32)), ((5, 6), ops.beam_splitter(6, 6, 0.2)), ((2,), ops.random_unitary(6, 33))]
    ref = psi
    for tg, op in layer:
        ref = oad.apply_operation_vector(dims, tg, ref, op)
    for group in (2, 3, 4):
        _lib.set_option("fuse_group", group)
        try:
            n0 = _lib.load().pw_launch_count()
            a, b = dev(psi).reshape(-1), torch.empty(N, dtype=torch.complex128, device="cuda")
            res, _ = ffi.circuit_device(a, b, False, dims, [(0, tg, dev(op)) for tg, op in layer], fuse=True)
            fused_launches = _lib.load().pw_launch_count() - n0
        finally:
            _lib.set_option("fuse_group", 3)
        assert rel_err(host(res).reshape(-1, 1), ref) < 1e-12, group
        assert fused_launches < len(layer) + 6


@pytest.mark.parametrize("cdtype", [np.complex128, np.complex64])
def test_dense_tensor_core_kernel_up_to_128_and_complex64(ad, cdtype):
    """64 < t <= 128 (32-fiber tiles, operator fragments from global memory) and the float2-storage
    instantiation of the dense FP64 tensor-core kernel; both fiber layouts, ragged tails."""
    from photon_weave_b200 import _lib

    tol = TOL[cdtype]
    _lib.set_option("blocks_min_elems", 0)
    try:
        for dims, tg, seed in [((64, 2, 70), (0, 1), 1), ((3, 128, 40), (1,), 2), ((77, 64, 2), (1, 2), 3),
                               ((5, 9, 11, 37), (1, 2), 4), ((33, 100), (1,), 5), ((6, 6, 50), (0, 1), 6),
                               ((50, 6, 6), (2, 1), 7), ((2, 8, 16, 19), (2, 1), 8)]:
            N = math.prod(dims)
            t = math.prod(dims[i] for i in tg)
            psi = ops.random_state(N, seed, cdtype)
            op = ops.random_unitary(t, 10 + seed).astype(cdtype)
            ref = oad.apply_operation_vector(dims, tg, psi.astype(np.complex128), op.astype(np.complex128), True)
            out = host(ad.apply_operation_vector(dims, tg, dev(psi), dev(op)))
            assert rel_err(out, ref) < tol, (dims, tg)
    finally:
        _lib.set_option("blocks_min_elems", 1 << 18)


def test_batched_jaynes_cummings_propagator_on_tensor_cores(ad):
    """examples/jaynes_cummings_model.py:130-141 for a BATCH of trajectories (the jax.vmap use of
    tests/state/test_shape_planning.py:104-121): B states of the cutoff-64 (x) qubit system under
    one 128 x 128 propagator = ONE launch of the dense tensor-core kernel (a GEMM 128x128 . 128xB)."""
    from photon_weave_b200 import _lib
    from photon_weave_b200.core import transforms as T

    dims, B = (64, 2), 3000
    U = ops.random_unitary(128, 5)
    states = np.stack([ops.random_state(128, 100 + i) for i in range(8)])
    batch = np.concatenate([states] * (B // 8))                       # (B, 128, 1)
    fn = lambda s, o: ad.apply_operation_vector(dims, (0, 1), s, o)
    _lib.set_option("blocks_min_elems", 0)
    try:
        n0 = _lib.load().pw_launch_count()
        out = T.vmap(fn, in_axes=(0, None))(dev(batch), dev(U))
        launches = _lib.load().pw_launch_count() - n0
    finally:
        _lib.set_option("blocks_min_elems", 1 << 18)
    want = np.stack([U @ s for s in states])
    got = host(out)
    assert got.shape == (B, 128, 1)
    assert rel_err(got[:8], want) < 1e-12 and rel_err(got[-8:], want) < 1e-12
    assert launches <= 4


def test_pinned_operator_plan_cache(ad):
    """pw_op_pin: the device-side analysis of a pinned operator is recorded once per (operator,
    dims, targets, dtype, aliasing) and replayed; results are bit-identical to the unpinned call
    on every structured path (blocks, staged blocks, two-sided, super-blocks, superoperator, dense,
    diagonal), across repeated calls, other streams, changed options and after unpinning."""
    from photon_weave_b200 import _lib
    from photon_weave_b200.core import ffi

    lib = _lib.load()
    _lib.set_option("blocks_min_elems", 0)
    _lib.set_option("blocks_min_elems_small_t", 0)
    try:
        cases = [("vec", (6, 6, 50), (0, 1), ops.beam_splitter(6, 6, 0.7)),
                 ("vec", (50, 6, 6), (1, 2), ops.beam_splitter(6, 6, 0.3)),
                 ("vec", (6, 6, 50), (0, 1), ops.random_unitary(36, 3)),
                 ("vec", (4, 6, 40), (1,), ops.phase_operator(6, 0.4)),
                 ("mat", (8, 4, 10), (0,), ops.random_unitary(8, 5)),
                 ("mat", (8, 8, 6), (0, 1), ops.beam_splitter(8, 8, 0.6)),
                 ("mat", (6, 8, 8), (1, 2), ops.beam_splitter(8, 8, 0.2)),
                 ("kraus", (5, 5, 12), (1,), ops.loss_channel(5, 0.1)),
                 ("kraus", (12, 5), (1,), ops.loss_channel(5, 0.2)),
                 ("kraus", (8, 10), (0,), np.stack([ops.random_unitary(8, 7 + k) for k in range(3)]) / np.sqrt(3))]
        skipped = 0
        for kind, dims, tg, op in cases:
            N = math.prod(dims)
            x = dev(ops.random_state(N, 1)).reshape(-1) if kind == "vec" else dev(ops.random_density(N, 2)).reshape(-1)
            fn = {"vec": ffi.apply_vec, "mat": ffi.apply_mat, "kraus": ffi.kraus_mat}[kind]
            plain = fn(x, dev(op), dims, tg)
            d_op = dev(op)
            before = lib.pw_plan_cache_size()
            ffi.pin_operator(d_op)
            first = fn(x, d_op, dims, tg)                       # records
            assert lib.pw_plan_cache_size() == before + 1
            n0 = lib.pw_launch_count()
            again = fn(x, d_op, dims, tg)                       # replays
            replay_launches = lib.pw_launch_count() - n0
            n0 = lib.pw_launch_count()
            fn(x, dev(op), dims, tg)
            plain_launches = lib.pw_launch_count() - n0
            assert replay_launches <= plain_launches, (kind, dims, tg)
            skipped += plain_launches - replay_launches                             # analysis kernels
            with torch.cuda.stream(torch.cuda.Stream()):
                other_stream = fn(x, d_op, dims, tg)
            torch.cuda.synchronize()
            x2 = x * (0.5 + 0.25j)                               # other data, same plan
            scaled = fn(x2, d_op, dims, tg)
            for got in (first, again, other_stream):
                assert torch.equal(got, plain), (kind, dims, tg)
            assert torch.equal(scaled, fn(x2, dev(op), dims, tg))
            _lib.set_option("seg_fpw", 0)                        # any option change: new epoch, re-recorded
            assert torch.equal(fn(x, d_op, dims, tg), plain)
            del d_op                                             # garbage collection unpins and frees
            import gc

            gc.collect()
            assert lib.pw_plan_cache_size() == before
        assert skipped >= 12
    finally:
        _lib.set_option("blocks_min_elems", 1 << 18)
        _lib.set_option("blocks_min_elems_small_t", 1 << 22)


def test_circuit_host_overlaps_upload_with_slab_wise_operations(ad):
    """pw_circuit_host on a state above the overlap threshold: operations that do not touch
    subsystem 0 (nor follow one that does on a shared subsystem) run slab by slab behind the
    upload; the result must equal the plain order on every kind of layer, including layers where
    NOTHING or EVERYTHING can run early, and match the non-overlapped call bit for bit."""
    from photon_weave_b200 import _lib
    from photon_weave_b200.core import ffi

    dims = (6,) * 7
    N = 6**7
    psi = ops.random_state(N, 71)
    u = [ops.random_unitary(6, 80 + m) for m in range(7)]
    bs = ops.beam_splitter(6, 6, 0.7)
    layers = {
        "bench_like": [((m,), u[m]) for m in range(7)] + [((0, 1), bs), ((2, 3), bs), ((4, 5), bs)],
        "late_first": [((0,), u[0]), ((0, 1), bs), ((1,), u[1]), ((3,), u[3]), ((1, 2), bs), ((5, 6), bs)],
        "all_early": [((m,), u[m]) for m in range(1, 7)] + [((3, 4), bs)],
        "all_late": [((0,), u[0]), ((0, 1), bs), ((1, 0), bs)],
    }
    for name, layer in layers.items():
        ref = psi
        for tg, op in layer:
            ref = oad.apply_operation_vector(dims, tg, ref, op)
        ops_list = [(0, tg, op) for tg, op in layer]
        outs = {}
        for thr in (1, 0):                      # 1 MiB threshold: overlapped; 0: plain
            _lib.set_option("host_overlap", thr)
            try:
                outs[thr] = ffi.circuit_host(psi.reshape(-1), False, dims, ops_list)
            finally:
                _lib.set_option("host_overlap", 256)
        assert rel_err(outs[1].reshape(-1, 1), ref) < 1e-12, name
        assert rel_err(outs[1], outs[0]) < 1e-13, name   # (different fusion groups: last-ulp differences)


def test_complex64_fiber_pairs_equal_single_fiber_kernels(ad):
    """complex64 register / block kernels: two adjacent fibers per thread (128-bit accesses) when the
    contiguous untouched run is even -- the same arithmetic in the same order, so results are
    bitwise those of the one-fiber kernels; odd runs and unaligned views keep the one-fiber path."""
    from photon_weave_b200 import _lib
    from photon_weave_b200.core import ffi

    cd = np.complex64
    cases = [((6, 5, 4), (0,), "dense"), ((3, 6, 7, 2), (1,), "dense"), ((8, 3, 10), (0,), "dense"),
             ((5, 5, 6), (0, 1), "bs"), ((4, 6, 6, 14), (1, 2), "bs"), ((6, 6, 9, 2), (0, 1), "bs"),
             ((7, 3, 4, 4), (0,), "phase"), ((6, 6, 5), (0, 1), "bs"), ((4, 7), (0,), "dense")]
    _lib.set_option("blocks_min_elems", 0)
    try:
        for dims, tg, kind in cases:
            N = math.prod(dims)
            t = math.prod(dims[i] for i in tg)
            if kind == "bs":
                op = ops.beam_splitter(dims[tg[0]], dims[tg[1]], 0.3 + 0.1 * len(dims))
            elif kind == "phase":
                op = ops.phase_operator(t, 0.7)
            else:
                op = ops.random_unitary(t, 3)
            op = np.ascontiguousarray(op).astype(cd)
            psi = ops.random_state(N, 7, cd)
            ref = oad.apply_operation_vector(dims, tg, psi.astype(np.complex128), op.astype(np.complex128), True)
            outs = []
            for pairs in (1, 0):
                _lib.set_option("c64_pairs", pairs)
                outs.append(host(ad.apply_operation_vector(dims, tg, dev(psi), dev(op))))
            assert np.array_equal(outs[0], outs[1]), (dims, tg, kind)
            assert rel_err(outs[0], ref) < TOL[cd], (dims, tg, kind)
        # a view that starts on an odd element: not 16-byte aligned -> one-fiber kernel, same result
        _lib.set_option("c64_pairs", 1)
        dims, tg = (6, 4, 6), (0,)
        N = math.prod(dims)
        op = np.ascontiguousarray(ops.random_unitary(6, 9)).astype(cd)
        psi = ops.random_state(N, 11, cd)
        buf = torch.zeros(N + 1, dtype=torch.complex64, device="cuda")
        buf[1:] = dev(psi).reshape(-1)
        res = ffi.apply_vec(buf[1:], dev(op), dims, tg)
        ref = oad.apply_operation_vector(dims, tg, psi.astype(np.complex128), op.astype(np.complex128), True)
        assert rel_err(host(res).reshape(-1, 1), ref) < TOL[cd]
    finally:
        _lib.set_option("c64_pairs", 1)
        _lib.set_option("blocks_min_elems", 1 << 18)


@pytest.mark.parametrize("cdtype", [np.complex128, np.complex64])
def test_row_parallel_kernels_for_launch_bound_states(ad, cdtype):
    """Dense operators on states below the analysis threshold: one block per fiber, or -- for a
    handful of fibers, the Jaynes-Cummings shape -- one warp per output row (k_apply_rows_warp)."""
    tol = TOL[cdtype]
    for dims, tg, seed in [((64, 2), (0, 1), 1), ((64, 2), (1, 0), 2), ((100,), (0,), 3), ((2, 40, 3), (1,), 4),
                           ((3, 33), (1,), 5), ((33, 3), (0,), 6), ((70, 9), (0,), 7), ((5, 12, 4, 3), (1, 3), 8),
                           ((200, 11), (0,), 9), ((9, 32), (1,), 10)]:
        N = math.prod(dims)
        t = math.prod(dims[i] for i in tg)
        psi = ops.random_state(N, seed, cdtype)
        op = ops.random_unitary(t, 20 + seed).astype(cdtype)
        ref = oad.apply_operation_vector(dims, tg, psi.astype(np.complex128), op.astype(np.complex128), True)
        out = host(ad.apply_operation_vector(dims, tg, dev(psi), dev(op)))
        assert rel_err(out, ref) < tol, (dims, tg)
