"""
CUDA-graph replay of a FIXED sequence of seam calls.

The launch-bound configurations of the reference (the Mach-Zehnder example: five calls on a 64-byte
state, ``examples/mach_zehnder_interferometer.py:36-58``; one Jaynes-Cummings step = apply +
partial trace, ``examples/jaynes_cummings_model.py:130-141``) cost a few microseconds of device
work and several kernel launches.  Every libpwb200 entry point only enqueues work on the current
stream, so such a sequence can be captured ONCE and replayed with a single graph launch:

    g = GraphedCalls(lambda psi: circuit(psi), psi0)     # warm-up + capture
    out = g(psi1)                                        # copy psi1 into the captured input, replay

The captured function must be shape-static and must not read results on the host (no ``.item()``,
no measurement outcome used as a Python int); operators should be pinned (`ffi.pin_operator`) so
that their structure analysis is replayed from the plan cache instead of being re-run.
"""

from __future__ import annotations

from typing import Any, Callable, Sequence

import torch


class GraphedCalls:
    """``fn(*inputs)`` captured into one CUDA graph; call the object to replay it.

    ``inputs`` are the STATIC input tensors of the capture: a replay with other tensors first
    copies them into those buffers (same shapes and dtypes).  The returned outputs are the
    graph's own output buffers -- overwritten by the next replay, so clone what must survive."""

    def __init__(self, fn: Callable[..., Any], *inputs: torch.Tensor, warmup: int = 3):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedCalls needs a CUDA device")
        self.inputs: Sequence[torch.Tensor] = inputs
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):      # first calls allocate plan-cache buffers: not capturable
            for _ in range(max(1, warmup)):
                fn(*inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn(*inputs)

    def __call__(self, *inputs: torch.Tensor):
        if len(inputs) not in (0, len(self.inputs)):
            raise ValueError(f"expected {len(self.inputs)} inputs (or none to reuse the captured ones)")
        for dst, src in zip(self.inputs, inputs):
            if src is not dst:
                if src.shape != dst.shape or src.dtype != dst.dtype:
                    raise ValueError("a graph replay needs inputs of the captured shapes and dtypes")
                dst.copy_(src)
        self.graph.replay()
        return self.outputs
