"""Key helpers of jax.random mapped onto the package's threefry (TEST INFRASTRUCTURE ONLY)."""

import numpy as _np

from photon_weave_b200.core import rng as _rng


def PRNGKey(seed):
    return _rng.PRNGKey(seed)


key = PRNGKey


def split(key, num=2):
    return _rng.split(key, num)


def _stream(key):
    k = _np.asarray(key).astype(_np.uint32).reshape(-1)
    return _np.random.Generator(_np.random.Philox(key=int(k[0]) << 32 | int(k[1])))


def poisson(key, lam, shape=()):
    return _stream(key).poisson(lam, size=shape)


def normal(key, shape=()):
    return _stream(key).normal(size=shape)
