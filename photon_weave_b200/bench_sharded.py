"""
bench.py --gpus N drivers: strong scaling of the 12-mode state-vector layer (configs[4]) and of
the 5-mode density-matrix Kraus / beam-splitter layer (configs[3]) over the ranks of one node,
plus the PARITY PREFLIGHT that runs a small state through the very same schedule on the same
ranks and compares the reassembled result with the oracle (the checker is injected by bench.py;
this package never imports ``oracle/``).
"""

from __future__ import annotations

import json
import os
from typing import Callable, List, Optional, Sequence

import torch

from .sharded import CudaBackend, PeerBuffers, PeerUnavailable, ShardedDensityMatrix, ShardedStateVector


def _device_ops(layer) -> List[torch.Tensor]:
    """Operators of a layer in HBM, pinned: they never change, so libpwb200 keeps their structure
    analysis across the many (chunk-wise) applications of the pipelined schedule."""
    from .core import ffi

    return [ffi.pin_operator(torch.as_tensor(o).cuda()) for _, o, _ in layer]


def make_peers(numel: int, nbuf: int, rank: int, world: int, dist, local_rank: int, want: bool):
    """(peers | None, description).  The decision is collective (PeerBuffers raises on EVERY rank
    or on none), so all ranks run the same swap implementation."""
    nccl = "torch.distributed.all_to_all_single (NCCL)"
    if not want or world == 1:
        return None, nccl if world > 1 else "single device (no swap)"
    try:
        peers = PeerBuffers(numel, torch.complex128, nbuf, rank, world, dist, local_rank)
    except PeerUnavailable as exc:
        return None, nccl + f" [peer memory unavailable: {exc}]"
    return peers, ("libpwb200 peer reads over NVLink (CUDA IPC): fused swap+apply kernels "
                   "(pw_apply_vec_gather), pw_gather_chunks otherwise")


class LayerRunner:
    """One circuit layer on a sharded state with TWO alternating layouts.

    mode "pipelined" (default): ``run_layer`` -- local operations that touch the outgoing axes,
    then the other local operations chunk by chunk with the axis swap hidden behind them (each
    finished chunk travels to its destination rank while the next one is computed), then the
    operations that needed the swap.
    mode "fused": everything local first, then the first operation that needs the swap carries
    it (fused gather + apply over peer memory, state vectors), then the rest.
    mode "plain": everything local, a stand-alone swap, the rest."""

    def __init__(self, st, layer: Sequence, d_ops: List[torch.Tensor], shard_a: Sequence[int],
                 shard_b: Sequence[int], apply: Callable, mode: str = "pipelined"):
        self.st, self.layer, self.d_ops = st, layer, d_ops
        self.shard_a, self.shard_b = list(shard_a), list(shard_b)
        self.apply = apply  # (st, op tensor, layer entry, prefer_shard)
        self.nops = len(layer)
        self.mode = mode
        self.last_info = None

    def order(self) -> List[int]:
        """Indices in the order the NEXT run applies them (for the oracle of the preflight the
        program order is used instead: regrouping must not change the result)."""
        return list(range(self.nops))

    def run(self, marks=None) -> int:
        st, layer, nops = self.st, self.layer, self.nops
        cur = set(st.sharded_axes())
        other = self.shard_b if cur == set(self.shard_a) else self.shard_a
        if self.mode in ("pipelined", "plain"):
            self.last_info = st.run_layer([(self.d_ops[i], layer[i][0]) for i in range(nops)],
                                          next_shard=other, pipeline=self.mode == "pipelined", marks=marks)
            return self.last_info["s0"] + self.last_info["s1"]

        def mark(label):
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((label, ev))

        first = [i for i in range(nops) if not set(layer[i][0]) & cur]
        second = [i for i in range(nops) if i not in first]
        mark("begin")
        for i in first:
            self.apply(st, self.d_ops[i], layer[i], None)
        mark("local_done")
        if second and (st.peers is None or not isinstance(st, ShardedStateVector)):
            st.ensure_local(tuple(a for i in second for a in layer[i][0]), prefer_shard=other)
        for n, i in enumerate(second):
            # peer mode (state vector): the first of these ops carries the swap (fused gather + apply)
            self.apply(st, self.d_ops[i], layer[i], other)
            if n == 0:
                mark("swap_done")
        mark("end")
        self.last_info = {"pipelined": False, "s0": 0, "s1": len(first), "s2": len(second)}
        return len(first)


def _apply_vec(st, op, entry, prefer):
    st.apply_operation(op, entry[0], prefer_shard=prefer)


def _apply_rho(st, op, entry, prefer):
    if op.dim() == 3:
        st.apply_kraus(op, entry[0], prefer_shard=prefer)
    else:
        st.apply_operation(op, entry[0], prefer_shard=prefer)


def _layouts(cls, dims, world, modes, layer=None):
    """Layout A = the initial sharding (leading subsystems); layout B = the window of as many
    subsystems that (i) shares no operation with A's and (ii) is touched by the FEWEST operations
    of the layer, dependency closure included: those operations (stage S0 of run_layer) cannot
    hide the swap, everything else local can."""
    vaxes, sparts = cls.initial_layout(dims, world)
    shard_a = sorted({s.g for s in sparts})
    k = max(len(shard_a), 1)
    mid = modes // 2
    best = list(range(mid, mid + k))
    if layer:
        targets = [set(t) for t, _, _ in layer]

        def cost(window):
            w = set(window)
            if w & set(shard_a) or any(t & w and t & set(shard_a) for t in targets):
                return None
            s0 = [bool(t & w) for t in targets]
            for j in reversed(range(len(targets))):          # operations that must precede an S0 one
                for i in range(j):
                    if s0[j] and not s0[i] and targets[i] & targets[j] and not targets[i] & set(shard_a):
                        s0[i] = True
            return sum(s0)

        # ties: the window nearest the front (sharding the innermost subsystems would shorten the
        # contiguous runs every local kernel streams)
        scored = [(cost(range(s, s + k)), s, list(range(s, s + k))) for s in range(k, modes - k + 1)]
        scored = [x for x in scored if x[0] is not None]
        if scored:
            best = min(scored)[2]
    return vaxes, sparts, shard_a, best


def parity_preflight(dist, rank: int, world: int, local_rank: int, kind: str, make_layer: Callable,
                     oracle_apply: Callable, random_input: Callable, want_peers: bool,
                     layers: int = 3, mode: str = "pipelined") -> dict:
    """A small state through the SAME LayerRunner schedule and swap implementation as the timed
    run, reassembled with gather_full() and compared with the oracle on rank 0.
    kind 'vec': 6^8 amplitudes; kind 'rho': a (4,4,4,4,4) density matrix (1024 x 1024)."""
    if kind == "vec":
        modes, cutoff = 8, 6
        dims = (cutoff,) * modes
        cls, apply = ShardedStateVector, _apply_vec
    else:
        modes, cutoff = 5, 4
        dims = (cutoff,) * modes
        cls, apply = ShardedDensityMatrix, _apply_rho
    layer = make_layer(modes, cutoff)
    full = random_input(kind, dims)                       # identical on every rank (seeded NumPy)
    vaxes, sparts, shard_a, shard_b = _layouts(cls, dims, world, modes, layer)
    own = cls._own_block(dims if kind == "vec" else dims + dims, torch.as_tensor(full).reshape(-1),
                         sparts, rank).cuda()
    peers, impl = make_peers(own.numel(), 3, rank, world, dist, local_rank, want_peers)
    spare = None
    if peers is not None:
        peers.tensors[0].copy_(own)
        own, spare = peers.tensors[0], peers.tensors[1]
    st = cls(dims, own, vaxes, sparts, rank, world, dist, CudaBackend(), spare=spare, peers=peers)
    d_ops = _device_ops(layer)
    runner = LayerRunner(st, layer, d_ops, shard_a, shard_b, apply, mode)
    for _ in range(layers):
        runner.run()
    got = st.gather_full().cpu().numpy()
    res = {"kind": kind, "dims": list(dims), "layers": layers, "ops": layers * len(layer),
           "swaps": st.swaps, "fused_swaps": st.fused_swaps, "pipelined_swaps": st.pipelined_swaps,
           "schedule": mode, "swap": impl.split(":")[0], "tolerance": 1e-12}
    if rank == 0:
        ref = full
        for _ in range(layers):                           # the oracle applies the PROGRAM order
            for tg, op, _ in layer:
                ref = oracle_apply(kind, dims, tg, ref, op)
        import numpy as np

        err = float(np.abs(got.reshape(ref.shape) - ref).max() / np.abs(ref).max())
        res.update(max_rel_err=err, ok=bool(err <= 1e-12))
    flag = torch.tensor([1.0 if (rank != 0 or res.get("ok")) else 0.0], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["ok"] = bool(flag.item() > 0)
    del st, runner, own, spare
    torch.cuda.synchronize()
    if peers is not None:
        if world > 1:
            dist.barrier()
        peers.close()
    return res


def _time_layers(args, dist, rank, world, local_rank, runner, st, lib, ClockSampler):
    """W warm-up layers, then exactly `steps` timed layers between barriers; returns the total
    (max over ranks) and the stage durations from the CUDA events run_layer recorded."""
    for _ in range(args.warmup):
        runner.run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    marks = [[] for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = lib.pw_launch_count()
    c0 = (st.swaps, st.packs, st.fused_swaps, st.pipelined_swaps)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    infos = []
    for s in range(args.steps):
        runner.run(marks[s])
        infos.append(dict(runner.last_info))
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = int(lib.pw_launch_count() - n0)
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([t0.elapsed_time(t1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    # stage durations: label -> ms per step (averaged over steps, max over ranks)
    names = ["s0", "s1", "swap_tail", "s2", "local", "swap"]
    acc = dict.fromkeys(names, 0.0)
    for m in marks:
        d = {lbl: ev for lbl, ev in m}

        def span(a, b):
            return d[a].elapsed_time(d[b]) if a in d and b in d else 0.0

        if "s1_done" in d:       # pipelined layer
            acc["s0"] += span("begin", "s0_done")
            acc["s1"] += span("s0_done", "s1_done")
            acc["swap_tail"] += span("s1_done", "swap_done")
            acc["s2"] += span("swap_done", "end")
        elif "swap_done" in d:   # plain / fused layer
            acc["local"] += span("begin", "local_done")
            acc["swap"] += span("local_done", "swap_done")
            acc["s2"] += span("swap_done", "end")
        else:
            acc["local"] += span("begin", "end")
    tm = torch.tensor([acc[n] / args.steps for n in names], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    stages = {n: float(v) for n, v in zip(names, tm)}
    counts = {"swaps": (st.swaps - c0[0]) / args.steps, "packs": (st.packs - c0[1]) / args.steps,
              "fused": (st.fused_swaps - c0[2]) / args.steps,
              "pipelined": (st.pipelined_swaps - c0[3]) / args.steps}
    return total_ms, stages, infos[-1], launches, clocks, counts


def _gpu_local_cpus(index: int):
    """CPUs on the NUMA node of GPU `index` (NVML), or None: pinned host memory is placed on the
    node of the allocating thread, and a shard crossing the socket link halves the PCIe rate."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if word >> b & 1}
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:
        return None


def _time_e2e(args, dist, rank, world, runner, st, units):
    """The step END TO END at N ranks: every rank uploads its shard from pinned host memory over
    its own PCIe link, runs the layer, and reads its shard back -- all inside the timed region
    (barrier + synchronize on both sides, max over ranks)."""
    steps = max(1, min(args.steps, 2 * getattr(args, "e2e_steps", 2)))
    before = os.sched_getaffinity(0)
    near = _gpu_local_cpus(torch.cuda.current_device())
    try:
        if near:
            os.sched_setaffinity(0, near)
    except OSError:      # not allowed in this container: allocate wherever the rank runs
        near = None
    try:
        host = torch.empty(st.local.numel(), dtype=st.local.dtype, pin_memory=True)
    finally:
        if near:
            os.sched_setaffinity(0, before)
    st.store_host(host)
    torch.cuda.synchronize()

    def step():
        st.load_host(host)
        runner.run()
        st.store_host(host)

    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([t0.elapsed_time(t1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / steps
    nbytes = host.numel() * host.element_size() * world
    del host
    return {"value": units / (ms_step * 1e-3), "ms_per_step": ms_step, "steps": steps,
            "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
            "host_numa": "pinned shard allocated on the GPU's NUMA node (NVML cpu affinity)" if near
            else "pinned shard allocated wherever the rank ran",
            "path": "per rank: pinned host shard -> ShardedTensor.load_host -> run_layer (axis swap "
                    "included) -> store_host; one PCIe link per rank, all ranks concurrently"}


def _records(world, n_local, nops, stages, info, counts, impl, mode, peak, peak_src):
    """roofline + axis_swap objects of the JSON line from the stage timings."""
    alg_bytes = 2 * n_local * 16
    swap_bytes = n_local * 16 * (world - 1) / world
    swap = {"swaps_per_step": counts["swaps"], "packs_per_step": counts["packs"],
            "fused_swaps_per_step": counts["fused"], "pipelined_swaps_per_step": counts["pipelined"],
            "schedule": mode, "bytes_received_per_rank": swap_bytes, "nvlink_reference_gbs": 770.0,
            "implementation": impl, "stage_ms": stages,
            "layer_split": {"s0_ops": info["s0"], "s1_ops": info["s1"], "s2_ops": info["s2"]}}
    if counts["pipelined"] > 0:
        # full-shard operations (S0 and S2) give the local kernel rate; S1 runs chunk by chunk
        # with the transfers behind it; what is left of the swap is swap_tail
        n_full = info["s0"] + info["s2"]
        avg_op_ms = (stages["s0"] + stages["s2"]) / max(n_full, 1)
        swap["exposed_swap_ms"] = stages["swap_tail"] + max(stages["s1"] - info["s1"] * avg_op_ms, 0.0)
        swap["s1_ms_per_op"] = stages["s1"] / max(info["s1"], 1)
        # the pieces move during S1: average inbound rate over that window
        swap["gbs_per_rank_inbound_during_s1"] = swap_bytes / ((stages["s1"] + stages["swap_tail"]) * 1e-3) / 1e9
        kernel = "apply (per-rank shard): full-shard operations of the layer (stages S0, S2)"
    elif counts["fused"] > 0:
        n_full = nops - 1
        avg_op_ms = (stages["local"] + stages["s2"]) / max(n_full, 1) if stages["swap"] > 0 else \
            (stages["local"] + stages["s2"]) / nops
        swap["fused_swap_plus_apply_ms"] = stages["swap"]
        swap["fused_gbs_per_rank_inbound"] = swap_bytes / (stages["swap"] * 1e-3) / 1e9 if stages["swap"] > 0 else None
        swap["exposed_swap_ms"] = stages["swap"] - avg_op_ms
        kernel = "apply (per-rank shard): local operations of the layer"
    else:
        avg_op_ms = (stages["local"] + stages["s2"]) / nops
        if counts["swaps"] > 0 and stages["swap"] > 0:
            swap["gbs_per_rank_per_direction"] = swap_bytes * counts["swaps"] / (stages["swap"] * 1e-3) / 1e9
        swap["exposed_swap_ms"] = stages["swap"]
        kernel = "apply (per-rank shard): all operations of the layer"
    achieved = alg_bytes / (avg_op_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_op_ms}
    return roofline, swap


def _schedule_mode(args) -> str:
    return getattr(args, "swap", None) or "pipelined"


def bench_sharded(args, dist, rank, world, local_rank, make_layer, workload_config, peaks,
                  ClockSampler, METRIC, UNIT, preflight: Optional[Callable] = None):
    """configs[4]: the 12-mode cutoff-6 state vector, strong scaling over `world` ranks."""
    from . import _lib
    from .core import ffi

    lib = _lib.load()
    want_peers = not getattr(args, "nccl_swap", False)
    mode = _schedule_mode(args)
    parity = preflight("vec", want_peers, mode) if preflight is not None else None
    modes, cutoff = args.modes, args.cutoff
    dims = (cutoff,) * modes
    N = cutoff**modes
    layer = make_layer(modes, cutoff)
    d_ops = _device_ops(layer)
    nops = len(layer)
    vaxes, sparts, shard_a, shard_b = _layouts(ShardedStateVector, dims, world, modes, layer)
    n_local = N // world
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    peers, swap_impl = make_peers(n_local, 3 if mode == "pipelined" else 2, rank, world, dist,
                                  local_rank, want_peers)
    if peers is not None:
        local, spare = peers.tensors[0], peers.tensors[1]
        torch.view_as_real(local).normal_(generator=gen)
    else:
        local = torch.view_as_complex(
            torch.randn(n_local, 2, generator=gen, device="cuda", dtype=torch.float64))
        spare = None
    st = ShardedStateVector(dims, local, vaxes, sparts, rank, world, dist, CudaBackend(),
                            spare=spare, peers=peers)
    nrm = st.norm2()
    ffi.scale_(st.local, (1.0 / torch.sqrt(nrm)).contiguous())
    runner = LayerRunner(st, layer, d_ops, shard_a, shard_b, _apply_vec, mode)
    total_ms, stages, info, launches, clocks, counts = _time_layers(
        args, dist, rank, world, local_rank, runner, st, lib, ClockSampler)
    norm_after = float(st.norm2().item())
    e2e = dict(_time_e2e(args, dist, rank, world, runner, st, nops * N), unit=UNIT)
    if rank == 0:
        peak, peak_src = peaks()
        roofline, swap = _records(world, n_local, nops, stages, info, counts, swap_impl, mode, peak, peak_src)
        line = {
            "metric": METRIC, "value": nops * N * args.steps / (total_ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": None, "axis_swap": swap,
            "norm_after": norm_after, "parity_n": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    if peers is not None:
        del st, local, spare, runner
        torch.cuda.synchronize()
        peers.close()
    if world > 1:
        dist.destroy_process_group()


RHO_METRIC = "density-matrix entries updated/sec (Kraus loss + beam-splitter layer, 5-mode cutoff-8 rho)"
RHO_UNIT = "entries/s"


def bench_sharded_rho(args, dist, rank, world, local_rank, make_rho_layer, peaks, ClockSampler,
                      preflight: Optional[Callable] = None):
    """configs[3]: the 5-mode cutoff-8 density matrix (17.18 GB), row-sharded over `world` ranks:
    a Kraus loss channel on every mode + beam splitters, one HBM pass per operation per block."""
    from . import _lib

    lib = _lib.load()
    want_peers = not getattr(args, "nccl_swap", False)
    mode = _schedule_mode(args)
    if mode == "fused":
        mode = "plain"   # there is no fused gather for the two-sided kernels
    parity = preflight("rho", want_peers, mode) if preflight is not None else None
    modes, cutoff = args.rho_modes, args.rho_cutoff
    dims = (cutoff,) * modes
    D = cutoff**modes
    layer = make_rho_layer(modes, cutoff)
    d_ops = _device_ops(layer)
    nops = len(layer)
    vaxes, sparts, shard_a, shard_b = _layouts(ShardedDensityMatrix, dims, world, modes, layer)
    n_local = D * D // world
    peers, swap_impl = make_peers(n_local, 3 if mode == "pipelined" else 2, rank, world, dist,
                                  local_rank, want_peers)
    # rho = A A^dagger / Tr with a (D, 4) Gaussian factor: PSD, built block by block
    gen = torch.Generator(device="cuda").manual_seed(4321)
    A = torch.view_as_complex(torch.randn(D, 4, 2, generator=gen, device="cuda", dtype=torch.float64))
    A = A / torch.linalg.matrix_norm(A)
    rows = D // world
    if peers is not None:
        local, spare = peers.tensors[0], peers.tensors[1]
    else:
        local, spare = torch.empty(n_local, dtype=torch.complex128, device="cuda"), None
    step = max(1, min(rows, (1 << 26) // D))
    lv = local.reshape(rows, D)
    for r0 in range(0, rows, step):
        torch.matmul(A[rank * rows + r0: rank * rows + r0 + step], A.conj().T, out=lv[r0:r0 + step])
    st = ShardedDensityMatrix(dims, local, vaxes, sparts, rank, world, dist, CudaBackend(),
                              spare=spare, peers=peers)
    tr0 = float(st.trace().item())
    runner = LayerRunner(st, layer, d_ops, shard_a, shard_b, _apply_rho, mode)
    total_ms, stages, info, launches, clocks, counts = _time_layers(
        args, dist, rank, world, local_rank, runner, st, lib, ClockSampler)
    tr1 = float(st.trace().item())
    e2e = dict(_time_e2e(args, dist, rank, world, runner, st, nops * D * D), unit=RHO_UNIT)
    if rank == 0:
        peak, peak_src = peaks()
        roofline, swap = _records(world, n_local, nops, stages, info, counts, swap_impl, mode, peak, peak_src)
        roofline["kernel"] = "Kraus / U rho U^dagger on a row block (one pass per op): " + roofline["kernel"]
        line = {
            "metric": RHO_METRIC, "value": nops * D * D * args.steps / (total_ms * 1e-3), "unit": RHO_UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": f"configs[3]: synthetic {modes}-mode density matrix, cutoff {cutoff} "
                                   f"({D * D} entries, {D * D * 16 / 1e9:.2f} GB complex128)",
                       "layer": ", ".join(lbl for _, _, lbl in layer), "ops_per_step": nops,
                       "sharding": "single device" if world == 1 else
                       f"leading ROW axis block-sharded over {world} ranks (column side local)",
                       "l2": "block (>= 2 GB per rank) exceeds the 126 MB L2; no flush needed"},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": None, "axis_swap": swap,
            "trace_before": tr0, "trace_after": tr1, "parity_n": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    if peers is not None:
        del st, local, spare, runner, lv
        torch.cuda.synchronize()
        peers.close()
    if world > 1:
        dist.destroy_process_group()
