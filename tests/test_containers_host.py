"""
Host logic of the state containers (routing, ordering, key threading, dimension
bookkeeping) with ``core.ffi`` routed to the NumPy oracle (tests/host_shim.py): runs without
a GPU.  tests/test_gpu_containers.py runs the same cases on the CUDA path.
"""

import pytest

import container_cases as cc
import host_shim
from photon_weave_b200.photon_weave import Config


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
def test_container_case_on_host_shim(case):
    with host_shim.host_backend():
        case()


def test_examples_on_host_shim():
    """examples/ against their analytic results (host logic; the CUDA run is in
    test_gpu_containers.py)."""
    import os
    import sys

    import numpy as np

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "examples"))
    import jaynes_cummings
    import mach_zehnder

    with host_shim.host_backend():
        pops, analytic, _, _ = jaynes_cummings.run(cutoff=8, steps=40)
        assert np.abs(pops - analytic).max() < 1e-9
        for phi, p, want, _ in mach_zehnder.run(points=5, shots=2):
            assert abs(p - want) < 1e-12
