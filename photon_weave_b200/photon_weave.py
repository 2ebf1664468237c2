"""
Config singleton + Session context manager (reference: photon_weave/photon_weave.py:10-137).
Keys are two uint32 words; splitting is host-side integer hashing (core/rng.py).
"""

from __future__ import annotations

import random
import sys
from typing import Any

import numpy as np

from .core import rng


class Config:
    _instance = None

    def __new__(cls, *args: Any, **kwargs: Any) -> "Config":
        if cls._instance is None:
            cls._instance = super(Config, cls).__new__(cls)
        return cls._instance

    def __init__(self) -> None:
        if not hasattr(self, "_initialized"):
            self._initialized = True
            self._random_seed = random.randint(0, sys.maxsize)
            self._key = rng.PRNGKey(self._random_seed)
            self._contractions = True
            self._dynamic_dimensions = False
            self._use_jit = False

    def set_seed(self, seed: int) -> None:
        """photon_weave.py:27-37."""
        self._random_seed = seed
        self._key = rng.PRNGKey(seed)

    @property
    def random_seed(self) -> int:
        return self._random_seed

    @property
    def random_key(self) -> np.ndarray:
        """Splits the current key and returns a new one (photon_weave.py:43-48)."""
        k = rng.split(self._key, 2)
        key, self._key = k[0], k[1]
        return key

    def set_contraction(self, cs: bool) -> None:
        self._contractions = cs

    @property
    def contractions(self) -> bool:
        return self._contractions

    @property
    def dynamic_dimensions(self) -> bool:
        return self._dynamic_dimensions

    @dynamic_dimensions.setter
    def dynamic_dimensions(self, v: bool) -> None:
        self._dynamic_dimensions = bool(v)

    def set_dynamic_dimensions(self, v: bool) -> None:
        self._dynamic_dimensions = bool(v)

    @property
    def use_jit(self) -> bool:
        return self._use_jit

    def set_use_jit(self, use_jit: bool) -> None:
        self._use_jit = bool(use_jit)


class Session:
    """Scope Config settings per run and restore them on exit (photon_weave.py:84-137)."""

    def __init__(self, *, seed=None, contractions=None, dynamic_dimensions=None, use_jit=None):
        cfg = Config()
        self._prev = {
            "seed": cfg.random_seed, "key": cfg._key, "contractions": cfg.contractions,
            "dynamic_dimensions": cfg.dynamic_dimensions, "use_jit": cfg.use_jit,
        }
        self._seed, self._contractions = seed, contractions
        self._dynamic_dimensions, self._use_jit = dynamic_dimensions, use_jit
        self._cfg = cfg

    def __enter__(self) -> Config:
        if self._seed is not None:
            self._cfg.set_seed(self._seed)
        if self._contractions is not None:
            self._cfg.set_contraction(self._contractions)
        if self._dynamic_dimensions is not None:
            self._cfg.set_dynamic_dimensions(self._dynamic_dimensions)
        if self._use_jit is not None:
            self._cfg.set_use_jit(self._use_jit)
        return self._cfg

    def __exit__(self, exc_type, exc, tb) -> None:
        self._cfg._random_seed = self._prev["seed"]
        self._cfg._key = self._prev["key"]
        self._cfg.set_contraction(self._prev["contractions"])
        self._cfg.set_dynamic_dimensions(self._prev["dynamic_dimensions"])
        self._cfg.set_use_jit(self._prev["use_jit"])
