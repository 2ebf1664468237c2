"""
The reference's L2 callers (state/utils/*.py, shape_planning.compiled_kernels, Config/
Session) restated on the CUDA seam, driven the way the reference's own tests drive them
(duck-typed DummyState objects).
"""

import numpy as np
import pytest

from oracle import adapters_np as oad
from oracle import operators as ops
from oracle import rng_threefry as orng

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


class DummyState:
    def __init__(self, dimensions):
        self.dimensions = dimensions


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


def host(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def basis(i, d):
    v = np.zeros((d, 1), dtype=np.complex128)
    v[i, 0] = 1
    return v


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as g

    g.build()


@pytest.fixture(autouse=True)
def restore_config():
    from photon_weave_b200.photon_weave import Config

    cfg = Config()
    prev = (cfg.contractions, cfg.dynamic_dimensions, cfg.use_jit)
    yield
    cfg.set_contraction(prev[0]); cfg.set_dynamic_dimensions(prev[1]); cfg.set_use_jit(prev[2])


def test_compiled_kernels_bundle():
    """tests/state/test_shape_planning.py:37-102."""
    from photon_weave_b200.state.utils.shape_planning import build_plan, compiled_kernels

    states = [DummyState(2), DummyState(2)]
    k = compiled_kernels(build_plan(states, [states[0]]))
    X = ops.x_operator()
    full = np.kron(X, np.eye(2))
    psi = basis(0, 4)
    assert np.allclose(host(k.apply_op_vector(dev(psi), dev(X))), full @ psi)
    rho = psi @ psi.conj().T
    assert np.allclose(host(k.apply_op_matrix(dev(rho), dev(X))), full @ rho @ full.conj().T)
    assert np.allclose(host(k.apply_kraus_matrix(dev(rho), dev(np.stack([X])))), full @ rho @ full.conj().T)
    o, post, _ = k.measure_vector(dev(psi), orng.PRNGKey(0))
    assert int(o) == 0 and np.allclose(host(post), basis(0, 2))
    st3 = [DummyState(2), DummyState(2), DummyState(2)]
    k3 = compiled_kernels(build_plan(st3, [st3[0], st3[2]]))
    rho8 = basis(0, 8) @ basis(0, 8).conj().T
    assert np.allclose(host(k3.trace_out_matrix(dev(rho8))), oad.trace_out_matrix((2, 2, 2), (0, 2), rho8))


def test_operations_bridge_noncontiguous_and_errors():
    """tests/state/utils/test_operations_contraction.py:32-108."""
    from photon_weave_b200.photon_weave import Config
    from photon_weave_b200.state.utils import operations
    from photon_weave_b200.state.utils.shape_planning import build_plan

    states = [DummyState(2), DummyState(3), DummyState(2)]
    targets = [states[2], states[0]]
    op = ops.random_unitary(4, 3)
    psi = ops.random_state(12, 4)
    plan = build_plan(states, targets)
    out = operations.apply_operation_vector(states, targets, dev(psi), dev(op), meta=plan)
    assert np.allclose(host(out), oad.apply_operation_vector((2, 3, 2), (2, 0), psi, op), atol=1e-13)
    rho = ops.random_density(12, 5)
    Config().set_contraction(True)
    out = operations.apply_operation_matrix(states, targets, dev(rho), dev(op))
    assert np.allclose(host(out), oad.apply_operation_matrix((2, 3, 2), (2, 0), rho, op), atol=1e-13)
    ks = [ops.random_unitary(4, 6) * 0.6, ops.random_unitary(4, 7) * 0.8]
    out = operations.apply_kraus_matrix(states, targets, dev(rho), [dev(k) for k in ks], meta=plan)
    assert np.allclose(host(out), oad.apply_kraus_matrix((2, 3, 2), (2, 0), rho, np.stack(ks)), atol=1e-13)
    with pytest.raises(AssertionError):
        operations.apply_operation_vector(states, targets, dev(psi), dev(np.eye(3)))
    with pytest.raises(AssertionError):
        operations.apply_kraus_matrix(states, targets, dev(rho), [dev(np.eye(3))])


def test_measurement_bridges_key_chains_match_oracle():
    """Key threading of measurements.py:50-80 (sequential) and :141-209 (joint + decode)."""
    from photon_weave_b200.state.utils import measurements as M

    states = [DummyState(3), DummyState(2), DummyState(2)]
    dims = (3, 2, 2)
    psi = ops.random_state(12, 8)
    rho = ops.random_density(12, 9)
    for seed in range(6):
        key = orng.PRNGKey(seed)
        # joint path: borrow_key, one draw over prod(target dims), mixed-radix decode
        outcomes, post, nxt = M.measure_vector_jit(states, [states[0], states[2]], dev(psi), key)
        use_key, next_key = orng.borrow_key(key)
        o_ref, post_ref, _ = oad.measure_vector(dims, (0, 2), psi, use_key)
        assert outcomes[states[0]] == o_ref // 2 and outcomes[states[2]] == o_ref % 2
        assert np.array_equal(np.asarray(nxt), next_key)
        assert np.allclose(host(post), post_ref, atol=1e-12)
        outcomes, post, _ = M.measure_matrix_jit(states, [states[1]], dev(rho), key)
        o_ref, post_ref, _ = oad.measure_matrix(dims, (1,), rho, use_key)
        assert outcomes[states[1]] == o_ref and np.allclose(host(post), post_ref, atol=1e-12)
        # sequential legacy path: keys = split(key, n+1)[1:], slices are not rescaled
        probs_seen = {}
        outcomes, post = M.measure_vector(states, [states[0], states[2]], dev(psi),
                                          prob_callback=lambda s, p: probs_seen.__setitem__(s, host(p)), key=key)
        keys = orng.split(key, 3)[1:]
        t = psi.reshape(3, 2, 2)
        p0 = np.sum(np.abs(t) ** 2, axis=(1, 2)); p0 /= p0.sum()
        o0 = orng.choice(keys[0], 3, p0)
        t = t[o0]
        p2 = np.sum(np.abs(t) ** 2, axis=0); p2 /= p2.sum()
        o2 = orng.choice(keys[1], 2, p2)
        assert outcomes[states[0]] == o0 and outcomes[states[2]] == o2
        assert np.allclose(probs_seen[states[0]], p0) and np.allclose(probs_seen[states[2]], p2)
        assert np.allclose(host(post).reshape(-1), t[:, o2], atol=1e-13)


def test_config_seed_kats_through_l2_layer():
    """tests/state/test_fock.py:346-380 through Config.random_key + measure_vector/matrix."""
    from photon_weave_b200.photon_weave import Config, Session
    from photon_weave_b200.state.utils import measurements as M

    f = DummyState(2)
    psi = np.array([[1], [1]], dtype=np.complex128) / np.sqrt(2)
    for seed, want in ((1, 0), (3, 1)):
        Config().set_seed(seed)
        out, _ = M.measure_vector([f], [f], dev(psi))
        assert out[f] == want
        Config().set_seed(seed)
        out, _ = M.measure_matrix([f], [f], dev(psi @ psi.conj().T))
        assert out[f] == want
    # POVM with PRNGKey(3) -> outcome 1 (tests/state/test_envelope.py:815-828)
    povm = [dev(np.diag([1.0, 0.0]).astype(np.complex128)), dev(np.diag([0.0, 1.0]).astype(np.complex128))]
    o, post = M.measure_POVM_matrix([f], [f], povm, dev(psi @ psi.conj().T), key=orng.PRNGKey(3))
    assert int(o) == 1 and np.allclose(host(post), np.diag([0.0, 1.0]))
    with Session(seed=7, contractions=False) as cfg:
        assert cfg.random_seed == 7 and cfg.contractions is False
    assert Config().contractions is True


def test_trace_out_bridges():
    """tests/state/utils/test_trace_out_contraction.py:18-45 and trace_out.py:17-25."""
    from photon_weave_b200.state.utils import trace_out as T

    states = [DummyState(2), DummyState(3), DummyState(2)]
    rho = ops.random_density(12, 11)
    psi = ops.random_state(12, 12)
    for uc in (False, True):
        out = T.trace_out_matrix(states, [states[2], states[0]], dev(rho), use_contraction=uc)
        assert np.allclose(host(out), oad.trace_out_matrix((2, 3, 2), (2, 0), rho, use_contraction=uc), atol=1e-13)
    out = T.trace_out_vector(states, [states[1]], dev(psi))
    assert np.allclose(host(out), oad.trace_out_vector((2, 3, 2), (1,), psi, True), atol=1e-13)
