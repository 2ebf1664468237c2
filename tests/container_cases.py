"""
Container-level cases (Fock / Polarization / CustomState / Envelope / CompositeEnvelope /
Operation) shared by ``test_containers_host.py`` (oracle-backed host shim, no GPU) and
``test_gpu_containers.py`` (the CUDA path).  Every expected value is either a golden of the
reference's own tests (cited per case, paths relative to /root/reference/tests) or is
recomputed with the NumPy oracle from the same inputs.

A case is a plain function; ``CASES`` lists them.  They only use the public API.
"""

from __future__ import annotations

import numpy as np

from oracle import adapters_np as oad
from oracle import operators as oops
from photon_weave_b200.core import rng
from photon_weave_b200.operation import (
    CompositeOperationType,
    CustomStateOperationType,
    FockOperationType,
    Operation,
    PolarizationOperationType,
)
from photon_weave_b200.photon_weave import Config
from photon_weave_b200.state.composite_envelope import CompositeEnvelope
from photon_weave_b200.state.custom_state import CustomState
from photon_weave_b200.state.envelope import Envelope
from photon_weave_b200.state.expansion_levels import ExpansionLevel
from photon_weave_b200.state.fock import Fock
from photon_weave_b200.state.polarization import Polarization, PolarizationLabel

S2 = 1 / np.sqrt(2)


def arr(x) -> np.ndarray:
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def close(x, expected, tol=1e-12) -> bool:
    x, e = arr(x), np.asarray(expected)
    return x.shape == e.shape and bool(np.allclose(x, e, atol=tol, rtol=tol))


def unit(dim: int, idx: int) -> np.ndarray:
    v = np.zeros((dim, 1))
    v[idx, 0] = 1
    return v


def last_kraus(ops):
    """The operator that completes sum K^dagger K = I (tests/state/test_composite_envelope.py,
    get_the_last_kraus_operator)."""
    d = ops[0].shape[0]
    rest = np.eye(d, dtype=complex) - sum(o.conj().T @ o for o in ops)
    w, v = np.linalg.eigh(rest)
    return v @ np.diag(np.sqrt(np.clip(w, 0, None))) @ v.conj().T


# ------------------------------------------------------------------------------------------
# single states
# ------------------------------------------------------------------------------------------


def fock_expand_contract_levels():
    """state/test_fock.py: label -> vector -> matrix -> back."""
    f = Fock()
    f.state = 2
    f.dimensions = 4
    f.expand()
    assert f.expansion_level == ExpansionLevel.Vector and close(f.state, unit(4, 2))
    f.expand()
    assert f.expansion_level == ExpansionLevel.Matrix
    assert close(f.state, unit(4, 2) @ unit(4, 2).T)
    f.contract(final=ExpansionLevel.Vector)
    assert f.expansion_level == ExpansionLevel.Vector and close(f.state, unit(4, 2))
    f.contract()
    assert f.expansion_level == ExpansionLevel.Label and f.state == 2


def fock_creation_ladder_dynamic():
    """operation/test_fock_operation.py:68-100: a^dagger twenty times, label/vector/matrix."""
    Config().set_dynamic_dimensions(True)
    op = Operation(FockOperationType.Creation)
    f = Fock()
    for i in range(20):
        f.apply_operation(op)
        assert f.state == i + 1
    Config().set_contraction(False)
    for levels in (1, 2):
        g = Fock()
        for _ in range(levels):
            g.expand()
        for i in range(6):
            g.apply_operation(op)
            a = arr(g.state)
            diag = np.abs(a[:, 0]) if levels == 1 else np.real(np.diag(a))
            assert np.argmax(diag) == i + 1 and np.isclose(diag.sum(), 1.0)


def fock_annihilate_vacuum_raises():
    """operation/test_fock_operation.py (annihilation of |0>): ValueError."""
    f = Fock()
    f.dimensions = 3
    f.expand()
    try:
        f.apply_operation(Operation(FockOperationType.Annihilation))
    except ValueError as e:
        assert "entirely composed of zeros" in str(e)
    else:
        raise AssertionError("annihilating the vacuum must raise")


def fock_phase_displace_squeeze_match_oracle():
    """Operators on a superposition, against oracle builders (_math/ops.py:230-333)."""
    Config().set_contraction(False)
    d = 8
    psi = oops.random_state(d, 5)
    for kind, kw, mat in (
        (FockOperationType.PhaseShift, {"phi": 0.3}, oops.phase_operator(d, 0.3)),
        (FockOperationType.Displace, {"alpha": 0.5}, oops.displacement_operator(d, 0.5)),
        (FockOperationType.Squeeze, {"zeta": 0.2 + 0.2j}, oops.squeezing_operator(d, 0.2 + 0.2j)),
    ):
        for is_matrix in (False, True):
            f = Fock()
            f.dimensions = d
            f.state = np.outer(psi, psi.conj()) if is_matrix else psi.reshape(-1, 1)
            f.expansion_level = ExpansionLevel.Matrix if is_matrix else ExpansionLevel.Vector
            op = Operation(kind, **kw)
            f.apply_operation(op)
            out = mat @ psi
            if is_matrix:
                ref = np.outer(out, out.conj())
                if op.renormalize:
                    ref = ref / np.linalg.norm(ref)
            else:
                ref = out.reshape(-1, 1)
                if op.renormalize:
                    ref = ref / np.linalg.norm(ref)
            assert close(f.state, ref, 1e-11), kind


def fock_measure_seed_kats():
    """state/test_fock.py:346-380: seed 1 -> 0, seed 3 -> 1 on (|0>+|1>)/sqrt2, vector and matrix."""
    for seed, want in ((1, 0), (3, 1)):
        for levels in (0, 1):
            Config().set_seed(seed)
            f = Fock()
            f.dimensions = 2
            f.expand()
            f.state = np.array([[S2], [S2]])
            for _ in range(levels):
                f.expand()
            out = f.measure()
            assert out[f] == want and f.measured


def fock_resize():
    """state/test_fock.py (resize): pad, refuse lossy truncation, allow lossless."""
    f = Fock()
    f.state = 2
    f.dimensions = 3
    f.expand()
    assert f.resize(5) and close(f.state, unit(5, 2))
    assert not f.resize(1)  # (the reference's bound is num_quanta < new + 1: fock.py:402)
    assert f.resize(3) and close(f.state, unit(3, 2))
    f.expand()
    assert f.resize(4) and arr(f.state).shape == (4, 4) and np.isclose(arr(f.state)[2, 2], 1)


def fock_pnr_ideal_detector():
    """state/test_fock.py:398-414: an ideal PNR detector returns the photon number and threads
    the key."""
    f = Fock()
    f.state = 3
    f.dimensions = 5
    key = rng.PRNGKey(7)
    outcomes, jitter, next_key = f.measure_pnr(key=key)
    assert outcomes[f] == 3 and np.all(np.asarray(jitter) == 0)
    assert next_key is not None and not np.array_equal(np.asarray(next_key), np.asarray(key))


def pnr_two_targets_in_reversed_order():
    """state/utils/measurements.py:781-783: the jitted PNR path measures and decodes in the
    GIVEN target order.  A basis state |2>_A |1>_B with targets (B, A) must report B=1, A=2
    (decoding in tensor order while zipping in given order would swap them)."""
    from photon_weave_b200.state.utils import measurements as ms

    class _Obj:
        def __init__(self, d):
            self.dimensions = d

    a, b = _Obj(4), _Obj(3)
    psi = np.zeros((12, 1), dtype=np.complex128)
    psi[2 * 3 + 1, 0] = 1.0
    for fn, st in ((ms.measure_pnr_vector, psi), (ms.measure_pnr_matrix, psi @ psi.conj().T)):
        counts, post, jitter, nk = fn([a, b], (b, a), st, key=rng.PRNGKey(5), use_jit=True)
        assert counts[b] == 1 and counts[a] == 2, counts
        counts, _, _, _ = fn([a, b], (a, b), st, key=rng.PRNGKey(5), use_jit=True)
        assert counts[a] == 2 and counts[b] == 1, counts


def polarization_labels_and_gates():
    """state/test_polarization.py + operation/test_polarization_operation.py."""
    for label, vec in ((PolarizationLabel.H, [1, 0]), (PolarizationLabel.V, [0, 1]),
                       (PolarizationLabel.R, [S2, 1j * S2]), (PolarizationLabel.L, [S2, -1j * S2]),
                       (PolarizationLabel.D, [S2, S2]), (PolarizationLabel.A, [S2, -S2])):
        p = Polarization(label)
        p.expand()
        assert close(p.state, np.array(vec).reshape(2, 1))
        p.expand()
        p.contract()
        assert p.state == label and p.expansion_level == ExpansionLevel.Label
    p = Polarization()
    p.apply_operation(Operation(PolarizationOperationType.X))
    assert p.state == PolarizationLabel.V
    p.apply_operation(Operation(PolarizationOperationType.H))
    assert p.state == PolarizationLabel.A
    Config().set_contraction(False)
    q = Polarization(PolarizationLabel.R)
    q.apply_operation(Operation(PolarizationOperationType.RX, theta=0.5))
    assert close(q.state, oops.rx_operator(0.5) @ np.array([[S2], [1j * S2]]))
    u = oops.random_unitary(2, 3)
    q.apply_operation(Operation(PolarizationOperationType.Custom, operator=u))
    assert close(q.state, u @ oops.rx_operator(0.5) @ np.array([[S2], [1j * S2]]))


def polarization_measure_seed_kats():
    """state/test_polarization.py:282-330: seed 1 -> H (0), seed 3 -> V (1) on |R>."""
    for seed, want in ((1, 0), (3, 1)):
        for levels in (1, 2):
            Config().set_seed(seed)
            p = Polarization(PolarizationLabel.R)
            for _ in range(levels):
                p.expand()
            assert p.measure()[p] == want


def custom_state_operation_kraus_povm():
    """state/test_custom_state.py, operation/test_custom_state_operation.py."""
    cs = CustomState(3)
    cs.expand()
    shift = np.roll(np.eye(3), 1, axis=0)
    cs.apply_operation(Operation(CustomStateOperationType.Custom, operator=shift))
    assert cs.state == 1
    ctx = {"a": lambda dims: np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]]),
           "b": lambda dims: np.eye(3)}
    cs2 = CustomState(3)
    cs2.apply_operation(Operation(CustomStateOperationType.Expresion,
                                  expr=("add", ("m_mult", "a", "a"), ("s_mult", 0, "b")), context=ctx))
    assert cs2.state == 2
    cs3 = CustomState(2)
    k = np.array([[0, 0], [1, 0]])
    cs3.apply_kraus([k, last_kraus([k])])
    assert cs3.state == 1
    cs4 = CustomState(2)
    cs4.state = np.array([[S2], [S2]])
    cs4.expansion_level = ExpansionLevel.Vector
    outcome, _ = cs4.measure_POVM([np.diag([1.0, 0.0]), np.diag([0.0, 1.0])], key=rng.PRNGKey(3))
    assert outcome == 1 and cs4.state == 1


# ------------------------------------------------------------------------------------------
# envelope
# ------------------------------------------------------------------------------------------


def envelope_combine_reorder_resize():
    """state/test_envelope.py (combine / reorder) and state/test_fock.py:440-510 (resize)."""
    env = Envelope()
    env.fock.state = 1
    env.fock.dimensions = 2
    env.polarization.state = PolarizationLabel.R
    env.combine()
    assert close(env.state, np.array([[0], [0], [S2], [1j * S2]]))
    assert env.fock.index == 0 and env.polarization.index == 1
    env.reorder(env.polarization, env.fock)
    assert close(env.state, np.array([[0], [S2], [0], [1j * S2]]))
    assert env.fock.index == 1 and env.polarization.index == 0
    assert env.resize_fock(3)
    assert close(env.state, np.array([[0], [S2], [0], [0], [1j * S2], [0]]))
    assert not env.resize_fock(1) and env.fock.index == 0  # the check leaves (fock, pol) order
    env.expand()
    env.reorder(env.polarization, env.fock)
    ref = np.array([[0], [S2], [0], [0], [1j * S2], [0]])
    assert close(env.state, ref @ ref.conj().T)


def envelope_measure_key_kats():
    """state/test_envelope.py:801-828: PRNGKey(3) -> fock 1, polarization 0; POVM outcome 1."""
    env = Envelope()
    env.fock.dimensions = 2
    env.fock.state = np.array([[1.0], [1.0]]) * S2
    env.fock.expansion_level = ExpansionLevel.Vector
    env.polarization.state = np.array([[1.0], [0.0]])
    env.polarization.expansion_level = ExpansionLevel.Vector
    env.combine()
    out = env.measure(key=rng.PRNGKey(3))
    assert out[env.fock] == 1 and out[env.polarization] == 0 and env.measured
    f = Fock()
    f.dimensions = 2
    f.state = np.array([[1.0], [1.0]]) * S2
    f.expansion_level = ExpansionLevel.Vector
    f.expand()
    outcome, _ = f.measure_POVM([np.diag([1.0, 0.0]), np.diag([0.0, 1.0])], destructive=False,
                                partial=True, key=rng.PRNGKey(3))
    assert outcome == 1


def envelope_operations_and_kraus():
    """operation/test_fock_operation.py:102-128 (creation on a combined envelope) and
    state/test_envelope.py (Kraus on the pair)."""
    for contraction in (False, True):
        Config().set_contraction(contraction)
        env = Envelope()
        env.fock.expand()
        env.polarization.expand()
        env.combine()
        op = Operation(FockOperationType.Creation)
        for i in range(5):
            env.apply_operation(op, env.fock)
            a = arr(env.state).reshape(-1)
            assert np.argmax(np.abs(a)) == (i + 1) * 2 and np.isclose(np.abs(a).sum(), 1)
        env.apply_operation(Operation(PolarizationOperationType.X), env.polarization)
        assert np.argmax(np.abs(arr(env.state).reshape(-1))) == 5 * 2 + 1
    Config().set_contraction(True)
    env = Envelope()
    env.fock.dimensions = 2
    k = np.zeros((4, 4))
    k[3, 0] = 1
    env.apply_kraus([k, last_kraus([k])], env.fock, env.polarization)
    assert close(env.state, unit(4, 3))
    red = env.trace_out(env.fock)
    assert close(red, np.diag([0.0, 1.0]))


# ------------------------------------------------------------------------------------------
# composite envelope
# ------------------------------------------------------------------------------------------


def composite_init_merges_containers():
    """state/test_composite_envelope.py:25-147."""
    e1, e2, e3 = Envelope(), Envelope(), Envelope()
    c1 = CompositeEnvelope(e1, e2)
    c2 = CompositeEnvelope(e2, e3)
    assert c1.uid == c2.uid
    assert set(c1.envelopes) == {e1, e2, e3} and len(c1.state_objs) == 6
    assert e1.composite_envelope.container is c2.container
    cs = CustomState(3)
    c3 = CompositeEnvelope(c2, cs)
    assert cs in c1.state_objs and c3.container is c1.container and c2.uid == c3.uid


def composite_combine_reorder_goldens():
    """state/test_composite_envelope.py:384-452 (vector / three states / matrix reordering)."""
    e1, e2 = Envelope(), Envelope()
    e2.polarization.state = PolarizationLabel.V
    ce = CompositeEnvelope(e1, e2)
    ce.combine(e1.polarization, e2.polarization)
    assert close(ce.product_states[0].state, unit(4, 1))
    assert e1.polarization.index == (0, 0) and e2.polarization.index == (0, 1)
    ce.reorder(e2.polarization, e1.polarization)
    assert close(ce.product_states[0].state, unit(4, 2))
    assert e1.polarization.index == (0, 1) and e2.polarization.index == (0, 0)
    ce.product_states[0].expand()
    ce.reorder(e1.polarization, e2.polarization)
    assert close(ce.product_states[0].state, unit(4, 1) @ unit(4, 1).T)

    e1, e2 = Envelope(), Envelope()
    e1.fock.dimensions = e2.fock.dimensions = 2
    e1.polarization.state = PolarizationLabel.R
    ce = CompositeEnvelope(e1, e2)
    ce.combine(e1.polarization, e1.fock, e2.fock)
    first = S2 * np.array([[1], [0], [0], [0], [1j], [0], [0], [0]])
    assert close(ce.product_states[0].state, first)
    ce.reorder(e1.fock, e2.fock, e1.polarization)
    assert (e1.fock.index, e2.fock.index, e1.polarization.index) == ((0, 0), (0, 1), (0, 2))
    assert close(ce.product_states[0].state, S2 * np.array([[1], [1j], [0], [0], [0], [0], [0], [0]]))


def composite_kraus_goldens():
    """state/test_composite_envelope.py:729-817 (incl. the resulting state_objs order)."""
    e1, e2, e3 = Envelope(), Envelope(), Envelope()
    for e in (e1, e2, e3):
        e.fock.dimensions = 2
    ce = CompositeEnvelope(e1, e2, e3)
    ce.combine(e1.fock, e2.fock, e3.fock)
    k = np.array([[0, 0], [1, 0]])
    ce.apply_kraus([k, last_kraus([k])], e3.fock)
    assert ce.product_states[0].state_objs == [e3.fock, e2.fock, e1.fock]
    assert close(ce.product_states[0].state, unit(8, 4))

    e1, e2 = Envelope(), Envelope()
    e1.fock.dimensions = e2.fock.dimensions = 2
    ce = CompositeEnvelope(e1, e2)
    ce.combine(e1.fock, e2.polarization)
    ce.combine(e1.polarization, e2.fock)
    k = np.zeros((4, 4))
    k[3, 0] = 1
    ce.apply_kraus([k, last_kraus([k])], e1.fock, e2.fock)
    assert close(ce.product_states[0].state, unit(16, 12))
    bad = np.zeros((3, 3))
    try:
        ce.apply_kraus([bad], e1.fock)
    except ValueError:
        pass
    else:
        raise AssertionError("wrong Kraus dimensions must raise")
    try:
        ce.apply_kraus([2 * np.eye(2)], e1.fock)
    except ValueError:
        pass
    else:
        raise AssertionError("a non trace-preserving channel must raise")


def composite_measure_after_reorder():
    """state/test_composite_envelope.py:454-475."""
    e1, e2 = Envelope(), Envelope()
    e1.fock.dimensions = e2.fock.dimensions = 2
    e1.fock.state, e2.fock.state = 0, 1
    ce = CompositeEnvelope(e1, e2)
    ce.combine(e1.fock, e2.fock)
    ce.reorder(e2.fock, e1.fock)
    out = ce.measure(e2.fock, e1.fock, destructive=False, key=rng.PRNGKey(0))
    assert out[e2.fock] == 1 and out[e1.fock] == 0
    assert len(ce.product_states) == 0 and e1.fock.index is None and e2.fock.index is None


def composite_beam_splitter_hom():
    """operation/test_composite_operation.py:15-71: BS(pi/4)|1,1> = i/sqrt2 (|0,2> + |2,0>),
    vector and (contracted) matrix form."""
    op = Operation(CompositeOperationType.NonPolarizingBeamSplitter, eta=np.pi / 4)
    for as_matrix in (False, True):
        e1, e2 = Envelope(), Envelope()
        e1.fock.state = e2.fock.state = 1
        ce = CompositeEnvelope(e1, e2)
        ce.combine(e1.fock, e2.fock)
        if as_matrix:
            ce.expand(e1.fock)
        ce.apply_operation(op, e1.fock, e2.fock)
        want = np.zeros((9, 1), dtype=complex)
        want[2] = want[6] = S2 if as_matrix else 1j * S2  # contraction fixes the global phase
        assert close(ce.product_states[0].state, want)
        assert e1.fock.dimensions == 3 and e2.fock.dimensions == 3


def composite_three_beam_splitters_probabilities():
    """state/utils/test_measurements.py:66-134: marginals 0.75 / 0.25 on all four modes."""
    from photon_weave_b200.state.utils.measurements import measure_matrix, measure_vector

    op = Operation(CompositeOperationType.NonPolarizingBeamSplitter, eta=np.pi / 4)
    for as_matrix in (False, True):
        envs = [Envelope() for _ in range(4)]
        envs[0].fock.state = 1
        ce = CompositeEnvelope(*envs)
        ce.apply_operation(op, envs[0].fock, envs[1].fock)
        ce.apply_operation(op, envs[0].fock, envs[2].fock)
        ce.apply_operation(op, envs[3].fock, envs[1].fock)
        focks = [e.fock for e in envs]
        ce.reorder(*focks)
        if as_matrix:
            envs[0].fock.expand()
        seen = {}
        fn = measure_matrix if as_matrix else measure_vector
        for f in focks:
            fn(focks, [f], ce.states[0].state, prob_callback=lambda s, p: seen.__setitem__(s, arr(p)))
        assert len(seen) == 4
        for p in seen.values():
            assert np.isclose(p[0], 0.75, atol=1e-9) and np.isclose(p[1], 0.25, atol=1e-9)


def composite_trace_out_goldens():
    """state/test_composite_envelope.py:1200-1395."""
    e1, e2 = Envelope(), Envelope()
    e1.fock.dimensions = e2.fock.dimensions = 2
    e1.fock.state = 1
    e2.polarization.state = PolarizationLabel.R
    ce = CompositeEnvelope(e1, e2)
    ce.combine(e1.fock, e2.polarization, e2.fock)
    red = ce.trace_out(e2.polarization)
    assert close(red, np.array([[0.5, -0.5j], [0.5j, 0.5]]))
    pair = ce.trace_out(e2.polarization, e1.fock)
    v = np.kron(np.array([[S2], [1j * S2]]), unit(2, 1))
    assert close(pair, v @ v.conj().T)
    assert ce.product_states[0].state_objs[:2] == [e2.polarization, e1.fock]
    ce.product_states[0].expand()
    assert close(ce.trace_out(e1.fock), np.diag([0.0, 1.0]))


def composite_povm():
    """state/test_composite_envelope.py:1019-1160."""
    e1, e2 = Envelope(), Envelope()
    e1.fock.dimensions = e2.fock.dimensions = 2
    e1.polarization.state = PolarizationLabel.V
    ce = CompositeEnvelope(e1, e2)
    ce.combine(e1.polarization, e2.polarization)
    ops = [np.diag([1.0, 0.0]), np.diag([0.0, 1.0])]
    outcome, _ = ce.measure_POVM(ops, e1.polarization, key=rng.PRNGKey(1))
    assert outcome == 1 and e1.polarization.measured
    assert ce.product_states[0].state_objs == [e2.polarization]
    try:
        ce.measure_POVM([np.eye(3)], e2.polarization)
    except ValueError:
        pass
    else:
        raise AssertionError("wrong POVM dimensions must raise")


def composite_dynamic_dimensions_ladder():
    """operation/test_fock_operation.py:130-160, 247-296: creation / annihilation inside a
    product state resize the Fock space on the fly."""
    Config().set_dynamic_dimensions(True)
    for contraction in (False, True):
        Config().set_contraction(contraction)
        e1 = Envelope()
        e1.fock.expand()
        e1.polarization.expand()
        e1.combine()
        e2 = Envelope()
        ce = CompositeEnvelope(e1, e2)
        ce.combine(e1.fock, e2.polarization)
        up = Operation(FockOperationType.Creation)
        for i in range(5):
            e1.apply_operation(up, e1.fock)
            want = np.zeros(((i + 2) * 2 * 2, 1))
            want[(i + 1) * 4, 0] = 1
            assert close(ce.product_states[0].state, want)
        down = Operation(FockOperationType.Annihilation)
        for i in range(5):
            e1.fock.apply_operation(down)
        try:
            e1.fock.apply_operation(down)
        except ValueError:
            pass
        else:
            raise AssertionError("annihilating the vacuum must raise")


def composite_expression_operation():
    """operation/test_composite_operation.py:74-91: an Expression operator across two spaces."""
    e1, e2 = Envelope(), Envelope()
    e1.fock.state = 1
    e1.fock.dimensions = e2.fock.dimensions = 3
    ce = CompositeEnvelope(e1, e2)
    ctx = {"a_dag": lambda d: oops.creation_operator(d[0]), "a": lambda d: oops.annihilation_operator(d[0]),
           "b_dag": lambda d: oops.creation_operator(d[1]), "b": lambda d: oops.annihilation_operator(d[1])}
    expr = ("expm", ("s_mult", 1j, np.pi / 4,
                     ("add", ("kron", "a", "b_dag"), ("kron", "a_dag", "b"))))
    op = Operation(CompositeOperationType.Expression, expr=expr, state_types=(Fock, Fock), context=ctx)
    ce.apply_operation(op, e1.fock, e2.fock)
    # shrinking each Fock space after the operation moves it to the front (reference behaviour)
    assert ce.product_states[0].state_objs == [e2.fock, e1.fock]
    ce.reorder(e1.fock, e2.fock)
    got = arr(ce.product_states[0].state)
    dims = [s.dimensions for s in ce.product_states[0].state_objs]
    assert dims == [2, 2]
    full = oops.beam_splitter(3, 3, np.pi / 4) @ np.kron(unit(3, 1), unit(3, 0))
    full = full.reshape(3, 3)[:dims[0], :dims[1]].reshape(-1, 1)
    assert close(got, full / np.linalg.norm(full), 1e-10)


def mach_zehnder_example():
    """examples/mach_zehnder_interferometer.py:36-58: BS, phase, BS; detection statistics follow
    cos^2 / sin^2 of half the phase."""
    for phi in (0.0, 0.7, np.pi / 2, np.pi):
        e1, e2 = Envelope(), Envelope()
        e1.fock.state = 1
        ce = CompositeEnvelope(e1, e2)
        bs = Operation(CompositeOperationType.NonPolarizingBeamSplitter, eta=np.pi / 4)
        ce.apply_operation(bs, e1.fock, e2.fock)
        e1.fock.apply_operation(Operation(FockOperationType.PhaseShift, phi=phi))
        ce.apply_operation(bs, e1.fock, e2.fock)
        probs, _ = ce.measure_expectation(e1.fock)
        p = arr(probs)
        assert np.isclose(p[1], np.sin(phi / 2) ** 2, atol=1e-10) and np.isclose(p.sum(), 1)


def lossy_circuit_small():
    """benchmarks/lossy_circuit (BASELINE configs[1] in miniature): beam splitters and a
    photon-loss Kraus channel on a 3-mode composite, against the oracle on the same rho."""
    Config().set_contraction(False)
    d = 3
    envs = [Envelope() for _ in range(3)]
    envs[0].fock.state = 1
    envs[1].fock.state = 1
    for e in envs:
        e.fock.dimensions = d
    ce = CompositeEnvelope(*envs)
    focks = [e.fock for e in envs]
    ce.combine(*focks)
    ps = ce.product_states[0]
    ps.expand()
    rho = arr(ps.state).copy()
    loss = oops.loss_channel(d, 0.1)
    bs = oops.beam_splitter(d, d, np.pi / 4)
    ce.apply_kraus(list(loss), focks[1])
    rho = oad.apply_kraus_matrix((d, d, d), (1,), rho, loss)
    assert [id(s) for s in ps.state_objs] == [id(focks[1]), id(focks[0]), id(focks[2])]
    ce.reorder(*focks)
    assert close(ps.state, rho, 1e-12)
    op = Operation(CompositeOperationType.NonPolarizingBeamSplitter, eta=np.pi / 4)
    ce.apply_operation(op, focks[2], focks[0])
    rho = oad.apply_operation_matrix((d, d, d), (2, 0), rho, bs)
    rho = rho / np.linalg.norm(rho)
    # every Fock space was cut back to its support (one photon per mode at most: d -> 2)
    dims = tuple(s.dimensions for s in focks)
    assert dims == (2, 2, 2)
    keep = np.ix_(*[range(k) for k in dims])
    cut = rho.reshape((d,) * 6)[np.ix_(*[range(k) for k in dims + dims])].reshape(8, 8)
    ce.reorder(*focks)
    assert close(ps.state, cut, 1e-12)
    assert close(ce.trace_out(focks[1]), oad.trace_out_matrix(dims, (1,), cut), 1e-12)


def lazy_layout_is_invisible():
    """The product state keeps its array in the physical order it was built in; every public
    read (state, trace_out, measurement) sees the logical order."""
    Config().set_contraction(False)
    dims = (2, 3, 4)
    states = [CustomState(d) for d in dims]
    psi = oops.random_state(24, 11)
    ce = CompositeEnvelope(*states)
    for s in states:
        s.expand()
    ce.combine(*states)
    ps = ce.product_states[0]
    ps.state = psi.reshape(-1, 1)
    ce.reorder(states[2], states[0])
    assert ps.state_objs == [states[2], states[0], states[1]]
    u = oops.random_unitary(4, 2)
    ce.apply_operation(Operation(CustomStateOperationType.Custom, operator=u), states[2])
    ref = oad.apply_operation_vector(dims, (2,), psi.reshape(-1, 1), u)
    ref = ref / np.linalg.norm(ref)
    assert close(ce.trace_out(states[1], states[2]), oad.trace_out_vector(dims, (1, 2), ref), 1e-12)
    # a multi-state trace_out puts the requested states first (reference behaviour)
    assert ps.state_objs == [states[1], states[2], states[0]]
    want = ref.reshape(dims).transpose(1, 2, 0).reshape(-1, 1)
    assert close(ps.state, want, 1e-12)
    assert states[1].index == (0, 0) and states[2].index == (0, 1) and states[0].index == (0, 2)
    out = ce.measure(states[0], key=rng.PRNGKey(5))
    assert out[states[0]] in (0, 1) and ps.state_objs == [states[1], states[2]]
    post = arr(ps.state).reshape(3, 4)
    sel = ref.reshape(dims)[out[states[0]]]
    assert close(post, sel / np.linalg.norm(sel), 1e-12)


CASES = [
    fock_expand_contract_levels,
    fock_creation_ladder_dynamic,
    fock_annihilate_vacuum_raises,
    fock_phase_displace_squeeze_match_oracle,
    fock_measure_seed_kats,
    fock_resize,
    fock_pnr_ideal_detector,
    pnr_two_targets_in_reversed_order,
    polarization_labels_and_gates,
    polarization_measure_seed_kats,
    custom_state_operation_kraus_povm,
    envelope_combine_reorder_resize,
    envelope_measure_key_kats,
    envelope_operations_and_kraus,
    composite_init_merges_containers,
    composite_combine_reorder_goldens,
    composite_kraus_goldens,
    composite_measure_after_reorder,
    composite_beam_splitter_hom,
    composite_three_beam_splitters_probabilities,
    composite_trace_out_goldens,
    composite_povm,
    composite_dynamic_dimensions_ladder,
    composite_expression_operation,
    mach_zehnder_example,
    lossy_circuit_small,
    lazy_layout_is_invisible,
]
