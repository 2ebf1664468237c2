#!/usr/bin/env python
"""Host<->device copy rates of this box with pinned buffers (what bounds bench.py's e2e):
H2D alone, D2H alone, both directions at once.  Prints one JSON object."""
import json

import torch

n = 8 << 30  # bytes per buffer
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


out = {"h2d_gbs": n / timed(h2d) / 1e9, "d2h_gbs": n / timed(d2h) / 1e9}
t = timed(both)
out["duplex_gbs_each_way"] = n / t / 1e9
out["bytes"] = n
print(json.dumps(out))
