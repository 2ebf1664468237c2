"""CUDA-graph replay of fixed seam-call sequences (photon_weave_b200/core/graphs.py): the replayed
results equal the eager ones and the oracle's, on the captured input and on fresh inputs."""
import math

import numpy as np
import pytest

from oracle import adapters_np as oad
from oracle import operators as ops

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(x):
    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm(a.reshape(-1) - b.reshape(-1)) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from photon_weave_b200.core import adapters, ffi, graphs

    return adapters, ffi, graphs


def test_mach_zehnder_sequence_as_one_graph(gpu):
    ad, ffi, graphs = gpu
    dims = (2, 2)
    bs, ph = ops.beam_splitter(2, 2, math.pi / 4), ops.phase_operator(2, 0.7)
    d_bs, d_ph = ffi.pin_operator(_dev(bs)), ffi.pin_operator(_dev(ph))

    def circuit(psi):
        x = ad.apply_operation_vector(dims, (0, 1), psi, d_bs)
        x = ad.apply_operation_vector(dims, (0,), x, d_ph)
        x = ad.apply_operation_vector(dims, (0, 1), x, d_bs)
        return x, ad.trace_out_vector(dims, (0,), x), ad.trace_out_vector(dims, (1,), x)

    def oracle(psi):
        x = oad.apply_operation_vector(dims, (0, 1), psi, bs, True)
        x = oad.apply_operation_vector(dims, (0,), x, ph, True)
        x = oad.apply_operation_vector(dims, (0, 1), x, bs, True)
        return x, oad.trace_out_vector(dims, (0,), x, True), oad.trace_out_vector(dims, (1,), x, True)

    psi0 = np.zeros((4, 1), dtype=np.complex128)
    psi0[2, 0] = 1
    g = graphs.GraphedCalls(circuit, _dev(psi0))
    for psi in (psi0, ops.random_state(4, 3), ops.random_state(4, 4)):
        outs = [o.cpu().numpy() for o in g(_dev(psi))]
        for got, ref in zip(outs, oracle(psi)):
            assert _rel(got, ref) < 1e-12
    with pytest.raises(ValueError):
        g(_dev(np.zeros((8, 1), dtype=np.complex128)))


def test_jaynes_cummings_step_and_lossy_rho_layer_as_graphs(gpu):
    ad, ffi, graphs = gpu
    from photon_weave_b200 import _lib

    # one Jaynes-Cummings step: dense 128 x 128 propagator + reduced qubit state
    dims = (64, 2)
    U = ops.random_unitary(128, 5)
    d_U = ffi.pin_operator(_dev(U))

    def step(psi):
        x = ad.apply_operation_vector(dims, (0, 1), psi, d_U)
        return x, ad.trace_out_vector(dims, (1,), x)

    psi = ops.random_state(128, 6)
    g = graphs.GraphedCalls(step, _dev(psi))
    x = psi
    for _ in range(3):   # chain three steps through the same graph
        out, red = g(_dev(x))
        x_ref = U @ x
        assert _rel(out.cpu().numpy(), x_ref) < 1e-12
        assert _rel(red.cpu().numpy(), oad.trace_out_vector(dims, (1,), x_ref, True)) < 1e-12
        x = x_ref
    # lossy-circuit layer on the 6.25 MB density matrix: beam splitter + loss channel, structure
    # analysis replayed from the plan cache (pinned operators)
    dims = (5, 5, 5, 5)
    bs5, loss = ops.beam_splitter(5, 5, math.pi / 4), ops.loss_channel(5, 0.1)
    d_bs5, d_loss = ffi.pin_operator(_dev(bs5)), ffi.pin_operator(_dev(loss))

    def layer(rho):
        r = ad.apply_operation_matrix(dims, (1, 2), rho, d_bs5)
        return ad.apply_kraus_matrix(dims, (2,), r, d_loss)

    rho = ops.random_density(625, 3)
    n0 = _lib.load().pw_launch_count()
    g2 = graphs.GraphedCalls(layer, _dev(rho))
    for seed in (3, 8):
        rho = ops.random_density(625, seed)
        got = g2(_dev(rho)).cpu().numpy()
        ref = oad.apply_kraus_matrix(dims, (2,), oad.apply_operation_matrix(dims, (1, 2), rho, bs5, True), loss, True)
        assert _rel(got, ref) < 1e-12
    assert _lib.load().pw_launch_count() > n0
