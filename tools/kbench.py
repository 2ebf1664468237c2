#!/usr/bin/env python
"""
Kernel micro-benchmark used while tuning (not the headline bench): times single
libpwb200 entry points on BASELINE-shaped states with CUDA events and prints
ms / algorithmic GB/s / fraction of the measured HBM peak per case.

    python tools/kbench.py vec   [--modes 12 --cutoff 6]     # configs[4] vector cases
    python tools/kbench.py mat   [--modes 5  --cutoff 8]     # configs[3] density-matrix cases
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
from oracle import operators as ops  # noqa: E402  (input generation only)
from photon_weave_b200 import _lib  # noqa: E402
from photon_weave_b200.core import ffi  # noqa: E402

PEAK = 6538.9
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


def report(name, ms, best, nbytes):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(f"{name:58s} {ms:9.3f} ms (min {best:9.3f})  {gbs:8.1f} GB/s  {gbs / PEAK:6.1%}", flush=True)


def rand_state(n, dtype=torch.complex128):
    gen = torch.Generator(device="cuda").manual_seed(0)
    rd = torch.float64 if dtype == torch.complex128 else torch.float32
    return torch.view_as_complex(torch.randn(n, 2, generator=gen, device="cuda", dtype=rd))


def dev(x, dtype=torch.complex128):
    return torch.as_tensor(np.ascontiguousarray(x)).to(dtype).cuda()


def bench_vec(args):
    modes, d = args.modes, args.cutoff
    dims = (d,) * modes
    N = d**modes
    dt = torch.complex64 if args.c64 else torch.complex128
    es = 8 if args.c64 else 16
    a = rand_state(N, dt)
    b = torch.empty_like(a)
    nbytes = 2 * N * es
    u = dev(ops.random_unitary(d, 1), dt)
    ph = dev(ops.phase_operator(d, 0.3), dt)
    bs = dev(ops.beam_splitter(d, d, math.pi / 4), dt)
    dense2 = dev(ops.random_unitary(d * d, 2), dt)
    axes = sorted(set([0, 1, modes // 2, modes - 3, modes - 2, modes - 1]))
    for path in args.paths:
        _lib.set_option("vec_path", path)
        for te in args.tile_elems:
            _lib.set_option("tile_elems", te)
            tag = f"[path={path} tile={te}]"
            for ax in axes:
                ms, best = timeit(lambda: ffi.apply_vec(a, u, dims, (ax,), out=b))
                report(f"{tag} dense {d}x{d} axis {ax}", ms, best, nbytes)
            ms, best = timeit(lambda: ffi.apply_vec(a, ph, dims, (modes // 2,), out=b))
            report(f"{tag} phase (diagonal) axis {modes // 2}", ms, best, nbytes)
            if path != 1 or args.slow:
                for pair in [(0, 1), (modes // 2, modes // 2 + 1), (modes - 2, modes - 1), (2, modes - 3),
                             (modes - 1, 0)]:
                    ms, best = timeit(lambda: ffi.apply_vec(a, bs, dims, pair, out=b), reps=3, warm=1)
                    report(f"{tag} beam splitter {pair}", ms, best, nbytes)
                ms, best = timeit(lambda: ffi.apply_vec(a, dense2, dims, (3, 4), out=b), reps=3, warm=1)
                report(f"{tag} dense {d*d}x{d*d} (3,4)", ms, best, nbytes)
    for pa, pb in [(3, 7), (0, 1), (modes - 2, modes - 1)]:
        r = ffi.apply_vec_pair(a, u, u, dims, pa, pb, out=b)
        if r is not None:
            ms, best = timeit(lambda: ffi.apply_vec_pair(a, u, u, dims, pa, pb, out=b))
            report(f"fused pair of {d}x{d} ops on axes ({pa},{pb}) [2 ops, 1 pass]", ms, best, nbytes)
    ms, best = timeit(lambda: ffi.probs(a, dims, (modes // 2,), False))
    report("probs_vec one mode", ms, best, N * es)
    ms, best = timeit(lambda: ffi.reduced_from_vec(a, dims, (modes // 2,)))
    report("reduced_from_vec one mode", ms, best, N * es)
    ms, best = timeit(lambda: ffi.norm_stats(a))
    report("norm_stats", ms, best, N * es)
    ms, best = timeit(lambda: ffi.occupancy(a, dims, False))
    report(f"occupancy (all {modes} modes' top level + norm, one pass)", ms, best, N * es)
    print(_lib.path_counters())


def bench_mat(args):
    modes, d = args.modes, args.cutoff
    dims = (d,) * modes
    D = d**modes
    dt = torch.complex64 if args.c64 else torch.complex128
    es = 8 if args.c64 else 16
    a = rand_state(D * D, dt)
    b = torch.empty_like(a)
    nbytes = 2 * D * D * es
    u = dev(ops.random_unitary(d, 1), dt)
    ph = dev(ops.phase_operator(d, 0.3), dt)
    loss = dev(ops.loss_channel(d, 0.1), dt)
    bs = dev(ops.beam_splitter(d, d, math.pi / 4), dt)
    axes = sorted(set([0, modes // 2, modes - 1]))
    for mode in args.paths:
        _lib.set_option("mat_mode", mode)
        for te in args.tile_elems:
            _lib.set_option("tile_elems", te)
            tag = f"[mat_mode={mode} tile={te}]"
            for ax in axes:
                ms, best = timeit(lambda: ffi.apply_mat(a, u, dims, (ax,), out=b), reps=3, warm=1)
                report(f"{tag} apply_mat dense {d}x{d} axis {ax}", ms, best, nbytes)
                ms, best = timeit(lambda: ffi.kraus_mat(a, loss, dims, (ax,), out=b), reps=3, warm=1)
                report(f"{tag} kraus loss K={d} axis {ax}", ms, best, nbytes)
            ms, best = timeit(lambda: ffi.apply_mat(a, ph, dims, (1,), out=b), reps=3, warm=1)
            report(f"{tag} apply_mat phase axis 1", ms, best, nbytes)
            if args.slow:
                for pair in [(0, 1), (1, 2), (modes - 2, modes - 1), (modes - 1, 0)]:
                    ms, best = timeit(lambda: ffi.apply_mat(a, bs, dims, pair, out=b), reps=3, warm=1)
                    report(f"{tag} apply_mat beam splitter {pair}", ms, best, nbytes)
    ms, best = timeit(lambda: ffi.probs(a.reshape(D, D), dims, (1,), True))
    report("probs_mat one mode (diagonal only)", ms, best, D * es)
    ms, best = timeit(lambda: ffi.trace_out_mat(a.reshape(D, D), dims, (1,)))
    report("trace_out_mat keep one mode", ms, best, d * d * (D // d) * es)
    print(_lib.path_counters())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["vec", "mat"])
    ap.add_argument("--modes", type=int, default=None)
    ap.add_argument("--cutoff", type=int, default=None)
    ap.add_argument("--paths", type=int, nargs="+", default=[0])
    ap.add_argument("--tile-elems", type=int, nargs="+", default=[0])
    ap.add_argument("--c64", action="store_true")
    ap.add_argument("--slow", action="store_true")
    ap.add_argument("--set", nargs="*", default=[], help="pw_set_option name=value ...")
    args = ap.parse_args()
    for kv in args.set:
        k, v = kv.split("=")
        _lib.set_option(k, int(v))
    print("options:", args.set, flush=True)
    if args.what == "vec":
        args.modes = args.modes or 12
        args.cutoff = args.cutoff or 6
        bench_vec(args)
    else:
        args.modes = args.modes or 5
        args.cutoff = args.cutoff or 8
        bench_mat(args)
