#!/usr/bin/env python
"""Host-side cost of the seam on launch-bound states (configs[0]: the Mach-Zehnder 5-call sequence):
wall time per call and the cProfile table of where the Python side spends it."""
import cProfile
import math
import os
import pstats
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402

g.build()
from oracle import operators as ops  # noqa: E402
from photon_weave_b200.core import adapters as ad  # noqa: E402
from photon_weave_b200.core import ffi  # noqa: E402

dims = (2, 2)
dev = lambda x: torch.as_tensor(np.ascontiguousarray(x)).cuda()
bs, ph = ffi.pin_operator(dev(ops.beam_splitter(2, 2, math.pi / 4))), ffi.pin_operator(dev(ops.phase_operator(2, 0.7)))
psi = np.zeros((4, 1), dtype=np.complex128)
psi[2, 0] = 1
d_psi = dev(psi)


def mzi():
    x = ad.apply_operation_vector(dims, (0, 1), d_psi, bs)
    x = ad.apply_operation_vector(dims, (0,), x, ph)
    x = ad.apply_operation_vector(dims, (0, 1), x, bs)
    return ad.trace_out_vector(dims, (0,), x), ad.trace_out_vector(dims, (1,), x)


def raw():
    x = ffi.apply_vec(d_psi, bs, dims, (0, 1))
    x = ffi.apply_vec(x, ph, dims, (0,))
    x = ffi.apply_vec(x, bs, dims, (0, 1))
    return ffi.reduced_from_vec(x, dims, (0,)), ffi.reduced_from_vec(x, dims, (1,))


jdims = (64, 2)
U = ffi.pin_operator(dev(ops.random_unitary(128, 5)))
jpsi = dev(ops.random_state(128, 6)).reshape(-1, 1)


def jc_step():
    x = ad.apply_operation_vector(jdims, (0, 1), jpsi, U)
    return ad.trace_out_vector(jdims, (1,), x)


for name, fn in (("adapters", mzi), ("ffi", raw), ("jc_step (2 calls)", jc_step)):
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000):
        fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 2000 * 1e6:.1f} us per sequence")
pr = cProfile.Profile()
pr.enable()
for _ in range(2000):
    mzi()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
