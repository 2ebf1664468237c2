"""
Measurement parity at LARGE joint outcome counts (t = 6^4, 6^6, 6^8): pw_draw's two tiers
(parallel locate with a rigorous error margin / one-warp sequential cumsum) against the
oracle's ``choice`` (core/kernels.py:204,224 call sites -> jax.random.choice), outcome
EQUALITY on hundreds of seeds, and the whole measure path (probs -> draw -> collapse,
core/kernels.py:189-227) against the oracle at the same sizes.
"""

import math

import numpy as np
import pytest

from oracle import adapters_np as oad
from oracle import operators as ops
from oracle import rng_threefry as orng

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ffi():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as g

    g.build()
    from photon_weave_b200.core import ffi as _ffi

    return _ffi


def _probs(t, seed, rdtype):
    rng = np.random.default_rng(seed)
    p = rng.random(t) ** 3  # uneven bins, some tiny
    p[rng.integers(0, t, size=max(t // 50, 1))] = 0.0  # empty bins
    p = (p / p.sum()).astype(rdtype)
    return p


@pytest.mark.parametrize("mode", [1, 2], ids=["two_tier", "forced_sequential"])
@pytest.mark.parametrize("rdtype", [np.float64, np.float32])
@pytest.mark.parametrize("t,seeds", [(6, 50), (100, 50), (4096, 50), (4097, 50), (6**4, 100),
                                     (6**6, 1000), (6**8, 100)])
def test_draw_equals_oracle_choice(ffi, t, seeds, rdtype, mode):
    from photon_weave_b200 import _lib

    if mode == 2 and t >= 6**8 and seeds > 20:
        seeds = 20  # every draw walks all 1.68 M bins on one warp
    if rdtype == np.float32 and t >= 6**6:
        seeds = min(seeds, 100)
    cdt = torch.complex128 if rdtype == np.float64 else torch.complex64
    p = _probs(t, 5 + t % 97, rdtype)
    d_p = torch.as_tensor(p).cuda()
    _lib.set_option("draw_exact", mode)
    try:
        got = [ffi.draw(d_p, orng.PRNGKey(s), cdt) for s in range(seeds)]
        got = torch.stack(got).cpu().numpy()
    finally:
        _lib.set_option("draw_exact", 1)
    want = np.array([orng.choice(orng.PRNGKey(s), t, p) for s in range(seeds)])
    want = np.minimum(want, t - 1)
    assert np.array_equal(got, want), (np.nonzero(got != want)[0][:5], got[:5], want[:5])


def test_draw_fast_tier_only_differs_within_rounding_of_an_edge(ffi):
    """draw_exact = 0 (no sequential tier): any difference from the oracle must be a draw that
    landed within the summation error of a bin edge."""
    from photon_weave_b200 import _lib

    t = 6**8
    p = _probs(t, 3, np.float64)
    d_p = torch.as_tensor(p).cuda()
    _lib.set_option("draw_exact", 0)
    try:
        got = torch.stack([ffi.draw(d_p, orng.PRNGKey(s), torch.complex128) for s in range(300)]).cpu().numpy()
    finally:
        _lib.set_option("draw_exact", 1)
    c = np.cumsum(p)
    for s in range(300):
        want = min(orng.choice(orng.PRNGKey(s), t, p), t - 1)
        if got[s] != want:
            r = c[-1] * (1.0 - orng.uniform(orng.PRNGKey(s), np.float64))
            lo, hi = min(got[s], want), max(got[s], want)
            assert np.abs(c[lo:hi] - r).max() < 4 * t * 1.2e-16, (s, got[s], want)


def test_draw_degenerate_inputs_stay_in_range(ffi):
    """All-zero / NaN probabilities (a zero state): the outcome is clamped to t - 1, the value
    an out-of-range gather index takes in JAX, instead of indexing past the arrays."""
    for t in (5, 5000):
        for fill in (0.0, float("nan")):
            d_p = torch.full((t,), fill, dtype=torch.float64, device="cuda")
            o = int(ffi.draw(d_p, orng.PRNGKey(1), torch.complex128))
            assert 0 <= o <= t - 1


@pytest.mark.parametrize("modes,targets", [(6, (4, 1, 3, 0)), (8, (7, 2, 0, 5, 3, 6)), (9, (1, 2, 3, 4, 5, 6, 7, 8))])
def test_measure_vector_large_joint_outcome(ffi, modes, targets):
    """probs -> draw -> collapse on 6^modes amplitudes with 6^4 / 6^6 / 6^8 joint outcomes.
    The device probabilities differ from NumPy's in the last ulp (reduction order), so an
    outcome may differ ONLY when the oracle's draw is within that noise of a bin edge."""
    from photon_weave_b200.core import adapters as ad

    dims = (6,) * modes
    N = 6**modes
    t = 6 ** len(targets)
    psi = ops.random_state(N, 40 + modes)
    d_psi = torch.as_tensor(psi).cuda()
    probs_ref = None
    nseeds = 100 if t <= 6**6 else 30
    mism = 0
    for seed in range(nseeds):
        key = orng.PRNGKey(1000 + seed)
        o, post, probs, _ = ad.measure_vector_with_probs(dims, targets, d_psi, key)
        if probs_ref is None:
            o_ref, post_ref, probs_ref, _ = oad.measure_vector_with_probs(dims, targets, psi, key)
            assert np.abs(probs.cpu().numpy() - probs_ref).max() <= 1e-12 * probs_ref.max()
            c = np.cumsum(probs_ref)
        else:
            o_ref = min(orng.choice(key, t, probs_ref), t - 1)
        if int(o) != int(o_ref):
            r = c[-1] * (1.0 - orng.uniform(key, np.float64))
            lo, hi = min(int(o), int(o_ref)), max(int(o), int(o_ref))
            assert np.abs(c[lo:hi] - r).max() < 1e-12, (seed, int(o), int(o_ref))
            mism += 1
            continue
        if seed < 3:
            rest = [i for i in range(modes) if i not in targets]
            sel = psi.reshape(dims).transpose(list(targets) + rest).reshape(t, -1)[int(o_ref)]
            want = sel / np.sqrt(probs_ref[int(o_ref)] + 1e-18)
            got = post.cpu().numpy().reshape(-1)
            assert np.abs(got - want).max() <= 1e-11 * max(np.abs(want).max(), 1e-30)
    assert mism <= 1


def test_measure_matrix_large_joint_outcome(ffi):
    """Density matrix, 6^4 joint outcomes of a 5-mode cutoff-6 rho (7776^2 entries)."""
    from photon_weave_b200.core import adapters as ad

    dims, targets = (6,) * 5, (3, 0, 4, 1)
    N, t = 6**5, 6**4
    rho = ops.random_density(N, 9, rank=3)
    d_rho = torch.as_tensor(rho).cuda()
    o_ref, post_ref, probs_ref, _ = oad.measure_matrix_with_probs(dims, targets, rho, orng.PRNGKey(0))
    c = np.cumsum(probs_ref)
    for seed in range(100):
        key = orng.PRNGKey(seed)
        o, post, probs, _ = ad.measure_matrix_with_probs(dims, targets, d_rho, key)
        want = min(orng.choice(key, t, probs_ref), t - 1)
        if seed == 0:
            assert np.abs(probs.cpu().numpy() - probs_ref).max() <= 1e-12 * probs_ref.max()
            assert int(o) == int(o_ref)
            assert np.abs(post.cpu().numpy() - post_ref).max() <= 1e-11 * np.abs(post_ref).max()
        if int(o) != want:
            r = c[-1] * (1.0 - orng.uniform(key, np.float64))
            lo, hi = min(int(o), want), max(int(o), want)
            assert np.abs(c[lo:hi] - r).max() < 1e-12, (seed, int(o), want)
