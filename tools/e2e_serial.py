#!/usr/bin/env python
"""pw_circuit_host latency on the 12-mode vector with and without the overlapped upload."""
import math, os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import __graft_entry__ as g
g.build()
from photon_weave_b200 import _lib
from photon_weave_b200.core import ffi
modes, cutoff = 12, 6
dims = (cutoff,) * modes
N = cutoff**modes
layer = bench.make_layer(modes, cutoff)
ops_list = [(0, t, o) for t, o, _ in layer]
h = torch.empty(N, dtype=torch.complex128, pin_memory=True)
h.real.normal_(); h.imag.normal_()
h /= h.norm()
hn = h.numpy()
for mode in (0, 256, 0, 256):
    _lib.set_option("host_overlap", mode)
    ffi.circuit_host(hn, False, dims, ops_list, out=hn)
    t0 = time.perf_counter()
    for _ in range(2):
        ffi.circuit_host(hn, False, dims, ops_list, out=hn)
    print("host_overlap=%d: %.1f ms per call" % (mode, (time.perf_counter() - t0) / 2 * 1e3), flush=True)
