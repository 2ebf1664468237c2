"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restatement of the JAX PRNG pieces the reference's measurement path calls
(third-party: jax 0.8.1, pinned in /root/reference/poetry.lock:507-508; the
source is NOT under /root/reference, so the published algorithm is restated):

* ``jax.random.PRNGKey(seed)``      (reference: photon_weave.py:22,37)
* ``jax.random.split(key, n)``      (reference: core/rng.py:37, photon_weave.py:47,
                                     state/utils/measurements.py:47,78)
* ``jax.random.choice(key, a, p=)`` (reference: core/kernels.py:204,224,243,261,341,
                                     state/utils/measurements.py:133,270)

The reference pins ``jax_default_prng_impl = "threefry2x32"``
(photon_weave/__init__.py:16) and enables x64 (_math/ops.py:11).  jax >= 0.5
defaults ``jax_threefry_partitionable=True``: ``split(key, n)[i]`` is
``threefry2x32(key, (hi=0, lo=i))`` and the 64 random bits of a scalar draw are
``hi<<32 | lo`` of ``threefry2x32(key, (0, 0))``.

Pinned by: the Random123 threefry2x32 known-answer vectors, the documented
``split(key(0))`` value, and the reference's own seed -> outcome KATs
(tests/state/test_fock.py:346-380, tests/state/test_polarization.py:282-330,
tests/state/test_envelope.py:801-828) -- see tests/test_oracle_rng.py.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np

_M32 = 0xFFFFFFFF
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x: int, r: int) -> int:
    return ((x << r) | (x >> (32 - r))) & _M32


def threefry2x32(k0: int, k1: int, x0: int, x1: int) -> Tuple[int, int]:
    """Threefry-2x32, 20 rounds (Salmon et al. 2011), as used by jax.prng."""
    ks = (k0 & _M32, k1 & _M32, (k0 ^ k1 ^ 0x1BD11BDA) & _M32)
    x0 = (x0 + ks[0]) & _M32
    x1 = (x1 + ks[1]) & _M32
    for i in range(5):
        for r in _ROT[i % 2]:
            x0 = (x0 + x1) & _M32
            x1 = _rotl(x1, r) ^ x0
        x0 = (x0 + ks[(i + 1) % 3]) & _M32
        x1 = (x1 + ks[(i + 2) % 3] + i + 1) & _M32
    return x0, x1


def PRNGKey(seed: int) -> np.ndarray:
    """jax.random.PRNGKey with x64 enabled: [seed >> 32, seed & 0xffffffff]."""
    seed = int(seed)
    return np.array([(seed >> 32) & _M32, seed & _M32], dtype=np.uint32)


def split(key, num: int = 2) -> np.ndarray:
    """jax.random.split (partitionable threefry): row i = threefry(key, (0, i))."""
    k0, k1 = int(key[0]), int(key[1])
    out = np.empty((num, 2), dtype=np.uint32)
    for i in range(num):
        out[i] = threefry2x32(k0, k1, (i >> 32) & _M32, i & _M32)
    return out


def random_bits64(key) -> int:
    a, b = threefry2x32(int(key[0]), int(key[1]), 0, 0)
    return (a << 32) | b


def random_bits32(key) -> int:
    a, b = threefry2x32(int(key[0]), int(key[1]), 0, 0)
    return a ^ b


def uniform(key, dtype=np.float64):
    """jax.random.uniform(key, (), dtype): mantissa-fill then subtract 1."""
    if np.dtype(dtype) == np.float64:
        bits = (random_bits64(key) >> 12) | 0x3FF0000000000000
        return np.array([bits], dtype=np.uint64).view(np.float64)[0] - np.float64(1.0)
    bits = (random_bits32(key) >> 9) | 0x3F800000
    return np.array([bits], dtype=np.uint32).view(np.float32)[0] - np.float32(1.0)


def choice(key, a: int, p) -> int:
    """
    jax.random.choice(key, a, p=p) with replace=True, shape=():
    ``c = cumsum(p); r = c[-1] * (1 - uniform(key)); searchsorted(c, r, 'left')``.
    """
    p = np.asarray(p)
    assert p.shape == (a,)
    c = np.cumsum(p)
    r = c[-1] * (c.dtype.type(1) - uniform(key, c.dtype))
    return int(np.searchsorted(c, r, side="left"))


def borrow_key(key):
    """core/rng.py:13-37."""
    if key is None:
        raise ValueError("PRNG key is required; got None")
    k = split(key)
    return k[0], k[1]
