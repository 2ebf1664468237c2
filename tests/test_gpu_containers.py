"""
The state containers on the CUDA path: the same cases as tests/test_containers_host.py
(goldens of the reference's container tests, oracle cross-checks), plus cases that push
large product states through the container API.
"""

import numpy as np
import pytest
import torch

import container_cases as cc
from photon_weave_b200.photon_weave import Config

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fresh_config():
    cfg = Config()
    prev = (cfg.contractions, cfg.dynamic_dimensions, cfg.use_jit)
    cfg.set_contraction(True)
    cfg.set_dynamic_dimensions(False)
    cfg.set_use_jit(False)
    cfg.set_seed(1234)
    yield
    cfg.set_contraction(prev[0])
    cfg.set_dynamic_dimensions(prev[1])
    cfg.set_use_jit(prev[2])


@pytest.mark.parametrize("case", cc.CASES, ids=lambda f: f.__name__)
def test_container_case_on_cuda(case):
    from photon_weave_b200 import _lib

    lib = _lib.load()
    before = lib.pw_launch_count()
    case()
    if case is not cc.composite_init_merges_containers:
        assert lib.pw_launch_count() > before, "the case did not launch any libpwb200 kernel"


def test_states_live_in_hbm():
    from photon_weave_b200.state.envelope import Envelope

    env = Envelope()
    env.fock.dimensions = 3
    env.combine()
    assert isinstance(env.state, torch.Tensor) and env.state.is_cuda
    assert env.state.dtype == torch.complex128


def test_large_product_state_through_the_container_api():
    """Eight cutoff-6 modes (1.7 M amplitudes) built, evolved, reordered, traced and measured
    through CompositeEnvelope, against the oracle on the same vector."""
    from oracle import adapters_np as oad
    from oracle import operators as oops
    from photon_weave_b200.operation import CompositeOperationType, FockOperationType, Operation
    from photon_weave_b200.state.composite_envelope import CompositeEnvelope
    from photon_weave_b200.state.envelope import Envelope

    Config().set_contraction(False)
    n, d = 8, 6
    dims = (d,) * n
    envs = [Envelope() for _ in range(n)]
    for e in envs:
        e.fock.dimensions = d
    focks = [e.fock for e in envs]
    ce = CompositeEnvelope(*envs)
    ce.combine(*focks)
    ps = ce.product_states[0]
    psi = oops.random_state(d**n, 3)
    ps.state = psi.reshape(-1, 1)
    ref = psi.reshape(-1, 1)

    def renorm(x):
        return x / np.linalg.norm(x)

    ce.apply_operation(Operation(FockOperationType.Displace, alpha=0.3), focks[5])
    ref = oad.apply_operation_vector(dims, (5,), ref, oops.displacement_operator(d, 0.3))
    ce.apply_operation(Operation(CompositeOperationType.NonPolarizingBeamSplitter, eta=np.pi / 4),
                       focks[6], focks[1])
    ref = renorm(oad.apply_operation_vector(dims, (6, 1), ref, oops.beam_splitter(d, d, np.pi / 4)))
    ce.apply_operation(Operation(FockOperationType.Squeeze, zeta=0.1 + 0.1j), focks[7])
    ref = renorm(oad.apply_operation_vector(dims, (7,), ref, oops.squeezing_operator(d, 0.1 + 0.1j)))
    assert [s.dimensions for s in focks] == [d] * n  # a dense state keeps every level occupied
    red = ce.trace_out(focks[7], focks[2])
    # (the oracle's trace_out_vector forms psi psi^dagger like the reference: too large here)
    t = ref.reshape(dims).transpose(7, 2, 0, 1, 3, 4, 5, 6).reshape(d * d, -1)
    want = t @ t.conj().T
    assert np.abs(red.cpu().numpy() - want).max() < 1e-12
    ce.reorder(*focks)
    got = ps.state.cpu().numpy()
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-12
    # expectation of measuring seven modes: joint probabilities + reduced state of the eighth
    probs, rest = ce.measure_expectation(*focks[:7])
    amp = ref.reshape(d**7, d)
    p = (np.abs(amp) ** 2).sum(axis=1)
    assert np.abs(probs.cpu().numpy() - p / p.sum()).max() < 1e-12
    assert np.abs(rest.cpu().numpy() - amp.T @ amp.conj()).max() < 1e-12


def test_examples_on_cuda():
    """examples/ (BASELINE configs[0] and configs[2] through the container API) against their
    analytic results."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "examples"))
    import jaynes_cummings
    import mach_zehnder

    pops, analytic, per_step, dims = jaynes_cummings.run(cutoff=64, steps=60)
    assert np.abs(pops - analytic).max() < 1e-9
    for phi, p, want, _ in mach_zehnder.run(points=5, shots=2):
        assert abs(p - want) < 1e-12
