"""
TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the parts of the ``jax`` API that the
REFERENCE'S OWN test-suite uses to build inputs and compare outputs (jax is not installed in
this image).  It lets /root/reference/tests run, unmodified, against photon_weave_b200 (see
run_reference_tests.py).  Nothing here computes anything on behalf of the product.
"""

import unittest

from . import numpy  # noqa: F401
from . import random  # noqa: F401
from . import lax  # noqa: F401


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


def jit(fn=None, **kwargs):
    if fn is None:
        return lambda f: f
    return fn


def _no_autodiff(*a, **k):
    raise unittest.SkipTest("autodiff needs the real jax (transform rules: SURVEY 8f.4)")


grad = value_and_grad = vmap = _no_autodiff
