"""
CPU-only checks of the drop-in boundary: libpwb200.so loads, exports every symbol
include/pw_b200.h declares, and its host-side integer entry points (no GPU work)
agree with the oracle.  Also checks the product never imports the oracle.
"""

import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from photon_weave_b200 import _lib

    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pw_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pw_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from photon_weave_b200 import _lib

    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in pw_b200.h but not exported"
    assert set(declared) == set(_lib.SIGNATURES), "ctypes table and header disagree"


def test_version_and_status_strings(lib):
    assert lib.pw_version() == 100
    assert lib.pw_status_string(0) == b"ok"
    assert b"invalid" in lib.pw_status_string(1)


def test_rng_split_matches_oracle_and_documented_value(lib):
    from oracle import rng_threefry as orng

    key = (ctypes.c_uint32 * 2)(0, 0)
    out = (ctypes.c_uint32 * 4)()
    assert lib.pw_rng_split(key, 2, out) == 0
    assert list(out) == [1797259609, 2579123966, 928981903, 3453687069]
    for seed in (1, 3, 123, 2**40 + 17):
        k = orng.PRNGKey(seed)
        kin = (ctypes.c_uint32 * 2)(int(k[0]), int(k[1]))
        buf = (ctypes.c_uint32 * 10)()
        assert lib.pw_rng_split(kin, 5, buf) == 0
        assert np.array_equal(np.array(buf).reshape(5, 2), orng.split(k, 5))


def test_invalid_arguments_return_status_not_crash(lib):
    dims = (ctypes.c_int64 * 2)(2, 2)
    tg = (ctypes.c_int32 * 1)(0)
    # null pointers are rejected before any CUDA call
    assert lib.pw_apply_vec(None, None, None, dims, 2, tg, 1, 1, None) == 1
    assert lib.pw_draw(None, 2, 0, 0, 1, None, None) == 1


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "photon_weave_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_no_cpu_fallback_without_a_cuda_device():
    """The product path fails loudly where there is no GPU (this container): no silent CPU route."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from photon_weave_b200.core import _arrays as A
    from photon_weave_b200.core.graphs import GraphedCalls

    with pytest.raises(RuntimeError):
        A.require_cuda()
    with pytest.raises(RuntimeError):
        GraphedCalls(lambda x: x, torch.zeros(2))
    from photon_weave_b200.core import adapters as ad
    import numpy as np

    with pytest.raises(RuntimeError):
        ad.apply_operation_vector((2,), (0,), np.array([[1.0], [0.0]]), np.eye(2))
