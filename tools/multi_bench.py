#!/usr/bin/env python
"""Times pw_apply_vec_multi (k fused single-mode operators per HBM pass) on the 12-mode cutoff-6
vector against k separate pw_apply_vec passes; prints a table (profiles/r2_multi_*.txt)."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402

g.build()
from oracle import operators as ops  # noqa: E402
from photon_weave_b200.core import ffi  # noqa: E402

modes = int(sys.argv[1]) if len(sys.argv) > 1 else 12
cdt = torch.complex128 if (len(sys.argv) < 3 or sys.argv[2] == "c128") else torch.complex64
d = 6
dims = (d,) * modes
N = d**modes
gen = torch.Generator(device="cuda").manual_seed(0)
rdt = torch.float64 if cdt == torch.complex128 else torch.float32
a = torch.view_as_complex(torch.randn(N, 2, generator=gen, device="cuda", dtype=rdt))
b = torch.empty_like(a)
us = [torch.as_tensor(np.ascontiguousarray(ops.random_unitary(d, i))).to(cdt).cuda() for i in range(4)]
es = 16 if cdt == torch.complex128 else 8


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


one = timed(lambda: ffi.apply_vec(a, us[0], dims, (5,), out=b))
print(f"# {modes}-mode cutoff-6 {cdt}, {N * es / 1e9:.2f} GB; one unfused pass (axis 5): {one:.2f} ms")
print("axes                 k   ms      ms/op   GB/s(alg, one pass)  vs k passes")
m = modes
cases = [(0, 1, 2), (3, 4, 5), (m - 6, m - 5, m - 4), (m - 3, m - 2, m - 1), (0, 5, m - 1), (m - 4, m - 3, m - 2),
         (0, 1), (m - 2, m - 1), (4, 5), (0, 1, 2, 3), (m - 4, m - 3, m - 2, m - 1), (2, 3, 4, 5)]
for axes in cases:
    k = len(axes)
    r = ffi.apply_vec_multi(a, us[:k], dims, axes, out=b)
    if r is None:
        print(f"{str(axes):20s} {k}   unsupported")
        continue
    ms = timed(lambda: ffi.apply_vec_multi(a, us[:k], dims, axes, out=b))
    print(f"{str(axes):20s} {k}   {ms:6.2f}  {ms / k:6.2f}  {2 * N * es / ms / 1e6:8.0f}             {k * one / ms:4.2f}x")
