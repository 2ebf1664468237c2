"""
Transform rules of the seam operations on the CUDA path (SURVEY.md section 8f.4,
photon_weave_b200/core/transforms.py):

* reverse-mode derivatives of apply / Kraus / expectation-mode measurement / partial trace
  against PyTorch's own autograd of the dense tensordot formula (a plain torch float64
  restatement of core/kernels.py:62-186, 28-59, 266-319), tolerance 1e-11;
* the reference's own gradient tests restated (tests/state/utils/test_measurements.py:337-371,
  tests/state/test_envelope.py:846-870: d/dtheta of p_0 = cos^2 theta is -sin 2 theta);
* the batching rule against the reference's vmap tests
  (tests/state/test_shape_planning.py:104-121, 149-153).
"""

import math

import numpy as np
import pytest

from oracle import operators as ops

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ad():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as g

    g.build()
    from photon_weave_b200.core import adapters

    return adapters


def dev(x, grad=False):
    t = torch.as_tensor(np.ascontiguousarray(x)).cuda()
    return t.requires_grad_(grad)


def dense_apply(psi, op, dims, targets):
    """(O on targets) psi with torch ops only (autograd oracle)."""
    n, nt = len(dims), len(targets)
    tdims = [dims[t] for t in targets]
    x = psi.reshape(dims)
    o = op.reshape(tdims + tdims)
    y = torch.tensordot(o, x, dims=(list(range(nt, 2 * nt)), list(targets)))
    return torch.movedim(y, list(range(nt)), list(targets)).reshape(-1, 1)


def dense_kraus(rho, kops, dims, targets):
    n = len(dims)
    N = math.prod(dims)
    out = 0
    for k in kops:
        y = dense_apply(rho.reshape(-1), k, tuple(dims) + tuple(dims), targets)
        y = dense_apply(y.reshape(-1), k.conj(), tuple(dims) + tuple(dims), [n + t for t in targets])
        out = out + y.reshape(N, N)
    return out


def real_loss(y, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    w = torch.randn(y.numel(), 2, generator=g, dtype=torch.float64).cuda()
    w = torch.view_as_complex(w).reshape(y.shape)
    return (w.conj() * y).real.sum() + (y.abs() ** 2).sum()


@pytest.mark.parametrize("dims,targets", [((3, 4, 2), (1,)), ((3, 4, 2), (2, 0)), ((6, 6, 6), (1, 2)), ((5,), (0,))])
def test_apply_vector_gradients(ad, dims, targets):
    N = math.prod(dims)
    t = math.prod(dims[i] for i in targets)
    psi0, op0 = ops.random_state(N, 1), ops.random_unitary(t, 2) * (1.0 + 0.3j)
    psi, op = dev(psi0, True), dev(op0, True)
    real_loss(ad.apply_operation_vector(dims, targets, psi, op), 3).backward()
    psi_r, op_r = dev(psi0, True), dev(op0, True)
    real_loss(dense_apply(psi_r, op_r, dims, targets), 3).backward()
    assert (psi.grad - psi_r.grad).abs().max() < 1e-11 * psi_r.grad.abs().max()
    assert (op.grad - op_r.grad).abs().max() < 1e-11 * op_r.grad.abs().max()


@pytest.mark.parametrize("dims,targets,K", [((3, 2, 2), (0,), 2), ((2, 3, 2), (2, 0), 1), ((4, 3), (1,), 3)])
def test_kraus_gradients(ad, dims, targets, K):
    N = math.prod(dims)
    t = math.prod(dims[i] for i in targets)
    rho0 = ops.random_density(N, 4) + 0.05 * ops.random_unitary(N, 5)   # not Hermitian on purpose
    k0 = np.stack([ops.random_unitary(t, 6 + k) * (0.7 - 0.2j * k) for k in range(K)])
    rho, ks = dev(rho0, True), dev(k0, True)
    real_loss(ad.apply_kraus_matrix(dims, targets, rho, ks), 7).backward()
    rho_r, ks_r = dev(rho0, True), dev(k0, True)
    real_loss(dense_kraus(rho_r, ks_r, dims, targets), 7).backward()
    assert (rho.grad - rho_r.grad).abs().max() < 1e-11 * rho_r.grad.abs().max()
    assert (ks.grad - ks_r.grad).abs().max() < 1e-11 * ks_r.grad.abs().max()
    if K == 1:
        rho2, u2 = dev(rho0, True), dev(k0[0], True)
        real_loss(ad.apply_operation_matrix(dims, targets, rho2, u2), 7).backward()
        assert (rho2.grad - rho_r.grad).abs().max() < 1e-11 * rho_r.grad.abs().max()
        assert (u2.grad - ks_r.grad[0]).abs().max() < 1e-11 * ks_r.grad.abs().max()


@pytest.mark.parametrize("dims,targets", [((2, 2), (0,)), ((3, 4, 2), (2, 0)), ((4, 3), (1,))])
def test_expectation_gradients_vector_and_matrix(ad, dims, targets):
    N = math.prod(dims)
    n = len(dims)
    rest = [i for i in range(n) if i not in targets]
    psi0 = ops.random_state(N, 8) * 1.3          # not normalised: the 1 / sum term matters
    psi = dev(psi0, True)
    p, post = ad.measure_vector_expectation(dims, targets, psi)
    (real_loss(p.to(torch.complex128), 9) + real_loss(post, 10)).backward()
    psi_r = dev(psi0, True)
    x = psi_r.reshape(dims).permute(list(targets) + rest).reshape(math.prod(dims[i] for i in targets), -1)
    s = (x.abs() ** 2).sum(dim=1)
    y = psi_r.reshape(dims).permute(rest + list(targets)).reshape(x.shape[1], -1)
    (real_loss((s / s.sum()).to(torch.complex128), 9) + real_loss(y @ y.conj().T, 10)).backward()
    assert (psi.grad - psi_r.grad).abs().max() < 1e-11 * psi_r.grad.abs().max()

    rho0 = ops.random_density(N, 11) * 1.7 + 0.02 * ops.random_unitary(N, 12)
    rho = dev(rho0, True)
    p, post = ad.measure_matrix_expectation(dims, targets, rho)
    (real_loss(p.to(torch.complex128), 13) + real_loss(post, 14)).backward()
    rho_r = dev(rho0, True)
    t4 = rho_r.reshape(tuple(dims) + tuple(dims))
    perm = list(targets) + rest
    m = t4.permute(perm + [n + q for q in perm]).reshape(x.shape[0], x.shape[1], x.shape[0], x.shape[1])
    s = torch.einsum("jrjr->j", m).real
    red = torch.einsum("jrjs->rs", m)
    (real_loss((s / s.sum()).to(torch.complex128), 13) + real_loss(red, 14)).backward()
    assert (rho.grad - rho_r.grad).abs().max() < 1e-11 * rho_r.grad.abs().max()


def test_reference_gradient_tests_restated(ad):
    """test_measurements.py:337-371 (finite gradient) and test_envelope.py:846-870
    (d p_0 / d theta = -sin 2 theta), vector and matrix pictures."""
    for theta0 in (0.5, 0.3):
        theta = torch.tensor(theta0, dtype=torch.float64, device="cuda", requires_grad=True)
        zero = torch.zeros((), dtype=torch.float64, device="cuda")
        state = torch.stack([torch.cos(theta), zero, torch.sin(theta), zero]).to(torch.complex128).reshape(4, 1)
        state = state / torch.linalg.vector_norm(state)
        probs, _ = ad.measure_vector_expectation((2, 2), (0,), state)
        (g,) = torch.autograd.grad(probs[0], theta)
        assert torch.isfinite(g) and abs(float(g) + math.sin(2 * theta0)) < 1e-10
        theta = torch.tensor(theta0, dtype=torch.float64, device="cuda", requires_grad=True)
        state = torch.stack([torch.cos(theta), zero, torch.sin(theta), zero]).to(torch.complex128).reshape(4, 1)
        rho = state @ state.conj().T
        probs, _ = ad.measure_matrix_expectation((2, 2), (0,), rho)
        (g,) = torch.autograd.grad(probs[0], theta)
        assert torch.isfinite(g) and abs(float(g) + math.sin(2 * theta0)) < 1e-10


def test_trace_out_gradients(ad):
    dims, keep = (3, 2, 4), (2, 0)
    N = math.prod(dims)
    rho0 = ops.random_density(N, 20) + 0.03 * ops.random_unitary(N, 21)
    rho = dev(rho0, True)
    real_loss(ad.trace_out_matrix(dims, keep, rho), 22).backward()
    rho_r = dev(rho0, True)
    t6 = rho_r.reshape(dims + dims)
    red = torch.einsum("abcdbf->cafd", t6).reshape(12, 12)   # keep (2, 0) in the given order
    real_loss(red, 22).backward()
    assert (rho.grad - rho_r.grad).abs().max() < 1e-11 * rho_r.grad.abs().max()
    psi0 = ops.random_state(N, 23)
    psi = dev(psi0, True)
    real_loss(ad.trace_out_vector(dims, keep, psi), 24).backward()
    psi_r = dev(psi0, True)
    y = psi_r.reshape(dims).permute(2, 0, 1).reshape(12, 2)
    real_loss(y @ y.conj().T, 24).backward()
    assert (psi.grad - psi_r.grad).abs().max() < 1e-11 * psi_r.grad.abs().max()


def test_vmap_batching_rule_matches_reference_tests(ad):
    """tests/state/test_shape_planning.py:104-121 and :149-153 restated."""
    from photon_weave_b200 import _lib
    from photon_weave_b200.core import transforms as T
    from photon_weave_b200.core.meta import make_meta
    from photon_weave_b200.state.utils.shape_planning import compiled_kernels

    lib = _lib.load()
    x = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    kernels = compiled_kernels(make_meta((2, 2), (0,)))
    batch = np.zeros((2, 4, 1), dtype=np.complex128)
    batch[0, 0, 0] = 1
    batch[1, 1, 0] = 1
    n0 = lib.pw_launch_count()
    out = T.vmap(kernels.apply_op_vector, in_axes=(0, None))(dev(batch), dev(x))
    assert lib.pw_launch_count() - n0 <= 2          # ONE apply for the whole batch
    full = np.kron(x, np.eye(2))
    assert np.allclose(out.cpu().numpy(), np.stack([full @ b for b in batch]))

    kernels = compiled_kernels(make_meta((2, 3, 2), (2,)))
    psi = np.zeros((12, 1), dtype=np.complex128)
    psi[0, 0] = 1
    rho = psi @ psi.conj().T
    rho2 = ops.random_density(12, 3)
    full = np.kron(np.eye(3), np.kron(np.eye(2), x))
    out = T.vmap(kernels.apply_kraus_matrix, in_axes=(0, None))(dev(np.stack([rho, rho2])), dev(x[None]))
    want = np.stack([full @ r @ full.conj().T for r in (rho, rho2)])
    assert np.abs(out.cpu().numpy() - want).max() < 1e-14
    # a bigger batch of states, dense operator, non-leading target; and the sequential fallback
    dims, tg = (3, 4, 2), (1,)
    op = ops.random_unitary(4, 5)
    states = np.stack([ops.random_state(24, 30 + i) for i in range(5)])
    fn = lambda s, o: ad.apply_operation_vector(dims, tg, s, o)
    a = T.vmap(fn, in_axes=(0, None))(dev(states), dev(op)).cpu().numpy()
    b = T.vmap(fn, in_axes=(0, None), method="sequential")(dev(states), dev(op)).cpu().numpy()
    from oracle import adapters_np as oad

    want = np.stack([oad.apply_operation_vector(dims, tg, s, op) for s in states])
    assert np.abs(a - want).max() < 1e-13 and np.abs(b - want).max() < 1e-13
    # sampling has no batching rule: vmap falls back to one call per element
    from oracle import rng_threefry as orng

    key = orng.PRNGKey(4)
    o, post, _ = T.vmap(lambda s: ad.measure_vector(dims, tg, s, key), in_axes=(0,))(dev(states))
    for i in range(5):
        o_ref, post_ref, _ = oad.measure_vector(dims, tg, states[i], key)
        assert int(o[i]) == int(o_ref)
