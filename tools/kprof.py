#!/usr/bin/env python
"""Runs ONE libpwb200 case a few times (for ncu captures).  Usage:
   python tools/kprof.py vec|mat <modes> <cutoff> <op: u|bs|phase|loss|krausd|dense2> <targets...> [--path P] [--tile E] [--reps R]"""
import argparse, math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.build()
from oracle import operators as ops
from photon_weave_b200 import _lib
from photon_weave_b200.core import ffi

ap = argparse.ArgumentParser()
ap.add_argument("what"); ap.add_argument("modes", type=int); ap.add_argument("cutoff", type=int)
ap.add_argument("op"); ap.add_argument("targets", type=int, nargs="+")
ap.add_argument("--path", type=int, default=0); ap.add_argument("--tile", type=int, default=0)
ap.add_argument("--ctas", type=int, default=0); ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--c64", action="store_true", help="complex64 state and operator")
ap.add_argument("--trace", action="store_true", help="print per-kernel device times (torch.profiler)")
ap.add_argument("--set", nargs="*", default=[], help="pw_set_option name=value ...")
a = ap.parse_args()
for kv in a.set:
    _k, _v = kv.split("=")
    _lib.set_option(_k, int(_v))
_lib.set_option("vec_path", a.path); _lib.set_option("mat_mode", a.path)
_lib.set_option("tile_elems", a.tile); _lib.set_option("ctas_per_sm", a.ctas)
d, m = a.cutoff, a.modes
dims = (d,) * m
N = d**m
n = N * N if a.what == "mat" else N
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.view_as_complex(torch.randn(n, 2, generator=gen, device="cuda",
                                      dtype=torch.float32 if a.c64 else torch.float64))
y = torch.empty_like(x)
t = d ** len(a.targets)
op = {"u": lambda: ops.random_unitary(t, 1), "bs": lambda: ops.beam_splitter(d, d, math.pi / 4),
      "phase": lambda: ops.phase_operator(d, 0.3), "loss": lambda: ops.loss_channel(d, 0.1),
      "dense2": lambda: ops.random_unitary(t, 2),
      "krausd": lambda: np.stack([ops.random_unitary(t, 10 + k) / 2.0 for k in range(4)])}[a.op]()
op = torch.as_tensor(np.ascontiguousarray(op)).cuda()
if a.c64:
    op = op.to(torch.complex64)
def run():
    for _ in range(a.reps):
        if a.what == "vec":
            ffi.apply_vec(x, op, dims, a.targets, out=y)
        elif a.op in ("loss", "krausd"):
            ffi.kraus_mat(x, op, dims, a.targets, out=y)
        else:
            ffi.apply_mat(x, op, dims, a.targets, out=y)
    torch.cuda.synchronize()


run()
if a.trace:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run()
    print(a.what, a.op, a.targets, a.set)
    for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
        print(f"  {ev.device_time_total / ev.count / 1e3:9.3f} ms x{ev.count:3d}  {ev.key[:110]}")
print("ok", _lib.path_counters())
