"""
Pins the oracle's PRNG restatement (oracle/rng_threefry.py) against

* the Random123 threefry2x32 known-answer vectors (the same three jax's own
  test-suite uses),
* the documented value of ``jax.random.split(jax.random.key(0))``,
* every seed -> outcome KAT the reference's tests hold for the measurement path.
"""

import numpy as np
import pytest

from oracle import adapters_np as ad
from oracle import rng_threefry as rng


@pytest.mark.parametrize(
    "key,ctr,expected",
    [
        ((0x0, 0x0), (0x0, 0x0), (0x6B200159, 0x99BA4EFE)),
        ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
        ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0)),
    ],
)
def test_threefry2x32_random123_kat(key, ctr, expected):
    assert rng.threefry2x32(*key, *ctr) == expected


def test_split_of_key0_matches_documented_value():
    k = rng.split(rng.PRNGKey(0))
    assert k.tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]


HALF = np.array([0.5, 0.5])


def test_config_seed_kats_fock_and_polarization():
    """
    tests/state/test_fock.py:346-380 and tests/state/test_polarization.py:282-330:
    Config.set_seed(1) -> outcome 0, set_seed(3) -> outcome 1 on a 50/50 state.
    Key chain: Config.random_key = split(PRNGKey(seed))[0] (photon_weave.py:47),
    then choice(key, arange(2), p) (state/utils/measurements.py:74-77,133).
    """
    assert rng.choice(rng.split(rng.PRNGKey(1))[0], 2, HALF) == 0
    assert rng.choice(rng.split(rng.PRNGKey(3))[0], 2, HALF) == 1


def test_envelope_measure_key3_kat():
    """
    tests/state/test_envelope.py:801-813: Envelope.measure(key=PRNGKey(3)) on
    (|0>+|1>)/sqrt2 (x) H gives fock 1, polarization 0.  Legacy sequential path:
    keys = split(key, 3)[1:] (measurements.py:78-80), one per target.
    """
    keys = rng.split(rng.PRNGKey(3), 3)[1:]
    psi = np.kron(np.array([1, 1]) / np.sqrt(2), np.array([1, 0])).reshape(4, 1)
    t = psi.reshape(2, 2)
    p_f = np.sum(np.abs(t) ** 2, axis=1)
    o_f = rng.choice(keys[0], 2, p_f / p_f.sum())
    assert o_f == 1
    rest = t[o_f]
    p_p = np.abs(rest) ** 2
    assert rng.choice(keys[1], 2, p_p / p_p.sum()) == 0


def test_povm_key3_kat():
    """tests/state/test_envelope.py:815-828: POVM outcome 1 with PRNGKey(3);
    use_key = split(key)[0] (measurements.py:637)."""
    psi = (np.array([[1.0], [1.0]]) / np.sqrt(2)).astype(np.complex128)
    rho = psi @ psi.conj().T
    ops = np.stack([np.diag([1.0, 0.0]), np.diag([0.0, 1.0])]).astype(np.complex128)
    use_key, _ = rng.borrow_key(rng.PRNGKey(3))
    outcome, post, _ = ad.measure_povm_matrix((2,), (0,), ops, rho, use_key)
    assert outcome == 1
    assert np.allclose(post, np.diag([0.0, 1.0]))


def test_keys_0_and_1_differ_key_123_reproducible():
    """tests/state/test_measurement_rng.py:21-39."""
    a = rng.choice(rng.split(rng.PRNGKey(0), 3)[1], 2, HALF)
    b = rng.choice(rng.split(rng.PRNGKey(1), 3)[1], 2, HALF)
    assert a != b
    c1 = rng.choice(rng.split(rng.PRNGKey(123), 3)[1], 2, HALF)
    c2 = rng.choice(rng.split(rng.PRNGKey(123), 3)[1], 2, HALF)
    assert c1 == c2


def test_deterministic_states_ignore_key():
    """tests/core/test_contraction_caching.py:64-83."""
    psi = np.zeros((2, 1), dtype=np.complex128)
    psi[1, 0] = 1
    o, post, _ = ad.measure_vector((2,), (0,), psi, rng.PRNGKey(0))
    assert o == 1 and post.shape == (1, 1)
    rho = np.diag([1.0, 0.0]).astype(np.complex128)
    o, post, _ = ad.measure_matrix((2,), (0,), rho, rng.PRNGKey(1))
    assert o == 0 and post.shape == (1, 1)


def test_borrow_key_requires_key():
    """core/rng.py:35-36."""
    with pytest.raises(ValueError):
        rng.borrow_key(None)
