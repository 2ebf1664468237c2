#!/usr/bin/env python
"""Measured arithmetic roofs of this B200 (FP64 FMA, FP64 tensor cores, FP32 FMA) through
pw_probe_arith; prints one JSON object (commit it as profiles/r2_arith_peaks.json)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402

g.build()
from photon_weave_b200 import _lib  # noqa: E402

lib = _lib.load()
sink = torch.zeros(1, dtype=torch.float64, device="cuda")
out = {}
KINDS = ((0, "fp64_fma"), (1, "fp64_mma_m8n8k4"), (2, "fp32_fma"), (3, "tf32_mma_sync_m16n8k8"),
         (4, "fp64_mma_with_1_dadd_per_6"), (5, "fp64_mma_with_2_dadd_per_6"), (6, "fp64_mma_with_3_dadd_per_6"),
         (7, "fp64_mma_with_2_fadd_per_6"))
for kind, name in KINDS:
    flops = C.c_double(0)
    best = 0.0
    for iters in (2000, 20000):
        _lib.check(lib.pw_probe_arith(kind, iters, sink.data_ptr(), C.byref(flops), None), "probe")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.pw_probe_arith(kind, iters, sink.data_ptr(), C.byref(flops),
                                      torch.cuda.current_stream().cuda_stream), "probe")
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    out[name + "_tflops"] = round(best, 2)
out["gpu"] = torch.cuda.get_device_name(0)
out["how"] = "register-only kernels, 148 x 8 CTAs x 256 threads, 16 independent chains per thread (8 MMAs per warp), CUDA events, best of two lengths"
print(json.dumps(out))
