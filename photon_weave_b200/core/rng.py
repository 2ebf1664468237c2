"""
Explicit PRNG key threading (reference: photon_weave/core/rng.py:13-37).

A key is two uint32 words (``jax.random.PRNGKey`` layout, x64 mode).  Splitting is
integer hashing done on the host by ``pw_rng_split`` (threefry2x32, partitionable
variant -- what jax 0.8.1 does by default); the random DRAW itself happens on the
device inside ``pw_draw``.
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import ffi


def PRNGKey(seed: int) -> np.ndarray:
    """jax.random.PRNGKey(seed) with jax_enable_x64 (reference: _math/ops.py:11)."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def split(key, num: int = 2) -> np.ndarray:
    """jax.random.split(key, num) -> (num, 2) uint32."""
    return ffi.rng_split(key, num)


def borrow_key(key: Optional[np.ndarray]) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """Return (use_key, next_key); requires an explicit key (core/rng.py:35-36)."""
    if key is None:
        raise ValueError("PRNG key is required; got None")
    k = split(key, 2)
    return k[0], k[1]
