"""
bench.py --gpus N drivers: strong scaling of the 12-mode state-vector layer (configs[4]) and of
the 5-mode density-matrix Kraus / beam-splitter layer (configs[3]) over the ranks of one node,
plus the PARITY PREFLIGHT that runs a small state through the very same schedule on the same
ranks and compares the reassembled result with the oracle (the checker is injected by bench.py;
this package never imports ``oracle/``).
"""

from __future__ import annotations

import json
import math
from typing import Any, Callable, List, Optional, Sequence

import torch

from .sharded import CudaBackend, PeerBuffers, PeerUnavailable, ShardedDensityMatrix, ShardedStateVector


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
    """One circuit layer on a sharded state with TWO alternating layouts: everything local under
    the current layout first, ONE axis swap (to the other layout), then the rest."""

    def __init__(self, st, layer: Sequence, d_ops: List[torch.Tensor], shard_a: Sequence[int],
                 shard_b: Sequence[int], apply: Callable):
        self.st, self.layer, self.d_ops = st, layer, d_ops
        self.shard_a, self.shard_b = list(shard_a), list(shard_b)
        self.apply = apply  # (st, op tensor, targets, prefer_shard, kind)
        self.nops = len(layer)

    def run(self, events=None) -> int:
        st, layer, nops = self.st, self.layer, self.nops
        cur = set(st.sharded_axes())
        first = [i for i in range(nops) if not set(layer[i][0]) & cur]
        second = [i for i in range(nops) if i not in first]
        other = self.shard_b if cur == set(self.shard_a) else self.shard_a
        slot = 0
        for i in first:
            if events is not None:
                events[slot].record(); slot += 1
            self.apply(st, self.d_ops[i], layer[i], None)
        if events is not None:
            events[slot].record(); slot += 1
        if second and (st.peers is None or not isinstance(st, ShardedStateVector)):
            # stand-alone swap (NCCL, or a plain peer gather for rho)
            st.ensure_local(tuple(a for i in second for a in layer[i][0]), prefer_shard=other)
        for i in second:
            if events is not None:
                events[slot].record(); slot += 1
            # peer mode (state vector): the first of these ops carries the swap (fused gather + apply)
            self.apply(st, self.d_ops[i], layer[i], other)
        if events is not None:
            events[slot].record()
        return len(first)


def _apply_vec(st, op, entry, prefer):
    st.apply_operation(op, entry[0], prefer_shard=prefer)


def _apply_rho(st, op, entry, prefer):
    if op.dim() == 3:
        st.apply_kraus(op, entry[0], prefer_shard=prefer)
    else:
        st.apply_operation(op, entry[0], prefer_shard=prefer)


def _layouts(cls, dims, world, modes):
    vaxes, sparts = cls.initial_layout(dims, world)
    shard_a = sorted({s.g for s in sparts})
    k = max(len(shard_a), 1)
    mid = modes // 2
    return vaxes, sparts, shard_a, list(range(mid, mid + k))


def parity_preflight(dist, rank: int, world: int, local_rank: int, kind: str, make_layer: Callable,
                     oracle_apply: Callable, random_input: Callable, want_peers: bool,
                     layers: int = 2) -> dict:
    """A small state through the SAME LayerRunner schedule and swap implementation as the timed
    run, reassembled with gather_full() and compared with the oracle on rank 0.
    kind 'vec': 6^8 amplitudes; kind 'rho': an (8,8,8) density matrix (512 x 512)."""
    if kind == "vec":
        modes, cutoff = 8, 6
        dims = (cutoff,) * modes
        cls, apply = ShardedStateVector, _apply_vec
    else:
        modes, cutoff = 3, 8
        dims = (cutoff,) * modes
        cls, apply = ShardedDensityMatrix, _apply_rho
    layer = make_layer(modes, cutoff)
    full = random_input(kind, dims)                       # identical on every rank (seeded NumPy)
    vaxes, sparts, shard_a, shard_b = _layouts(cls, dims, world, modes)
    own = cls._own_block(dims if kind == "vec" else dims + dims, torch.as_tensor(full).reshape(-1),
                         sparts, rank).cuda()
    peers, impl = make_peers(own.numel(), 2, rank, world, dist, local_rank, want_peers)
    spare = None
    if peers is not None:
        peers.tensors[0].copy_(own)
        own, spare = peers.tensors[0], peers.tensors[1]
    st = cls(dims, own, vaxes, sparts, rank, world, dist, CudaBackend(), spare=spare, peers=peers)
    d_ops = [torch.as_tensor(o).cuda() for _, o, _ in layer]
    runner = LayerRunner(st, layer, d_ops, shard_a, shard_b, apply)
    order: List[int] = []
    for _ in range(layers):
        cur = set(st.sharded_axes())
        first = [i for i in range(len(layer)) if not set(layer[i][0]) & cur]
        order += first + [i for i in range(len(layer)) if i not in first]
        runner.run()
    got = st.gather_full().cpu().numpy()
    swaps, fused = st.swaps, st.fused_swaps
    res = {"kind": kind, "dims": list(dims), "layers": layers, "ops": len(order), "swaps": swaps,
           "fused_swaps": fused, "swap": impl.split(":")[0], "tolerance": 1e-12}
    if rank == 0:
        ref = full
        for i in order:                                   # the order the schedule applied them in
            ref = oracle_apply(kind, dims, layer[i][0], ref, layer[i][1])
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


def _time_layers(args, dist, rank, world, local_rank, runner, st, nops, lib, ClockSampler):
    for _ in range(args.warmup):
        runner.run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(nops + 2)] for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = lib.pw_launch_count()
    c0 = (st.swaps, st.packs, st.fused_swaps)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    nfirst = [runner.run(ev[s]) for s in range(args.steps)]
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
    swap_ms = op_ms = fused_op_ms = 0.0
    for s in range(args.steps):
        nf = nfirst[s]
        for j in range(nops + 1):
            dt = ev[s][j].elapsed_time(ev[s][j + 1])
            if j == nf:
                swap_ms += dt
            else:
                op_ms += dt
            if j == nf + 1:
                fused_op_ms += dt  # first op after the swap slot (carries the fused swap)
    tm = torch.tensor([swap_ms / args.steps, op_ms / args.steps, fused_op_ms / args.steps],
                      device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    counts = ((st.swaps - c0[0]) / args.steps, (st.packs - c0[1]) / args.steps,
              (st.fused_swaps - c0[2]) / args.steps)
    return total_ms, float(tm[0]), float(tm[1]), float(tm[2]), launches, clocks, counts


def _swap_record(world, n_local, swap_ms, fused_op_ms, avg_op_ms, counts, impl, peers):
    swaps, packs, fused = counts
    swap_bytes = n_local * 16 * (world - 1) / world
    rec = {"swaps_per_step": swaps, "packs_per_step": packs, "ms_per_step": swap_ms,
           "bytes_received_per_rank": swap_bytes, "nvlink_reference_gbs": 770.0,
           "fused_swaps_per_step": fused, "implementation": impl,
           "gbs_per_rank_per_direction": None, "fused_swap_plus_apply_ms": None}
    if swaps - fused > 0 and swap_ms > 0:        # stand-alone swaps have a duration of their own
        rec["gbs_per_rank_per_direction"] = swap_bytes * (swaps - fused) / (swap_ms * 1e-3) / 1e9
    if fused > 0 and peers is not None:
        # the fused kernel streams (W-1)/W of the shard in over NVLink while it applies the operator
        rec["fused_swap_plus_apply_ms"] = fused_op_ms
        rec["fused_gbs_per_rank_inbound"] = swap_bytes / (fused_op_ms * 1e-3) / 1e9
        rec["fused_overhead_vs_local_op_ms"] = fused_op_ms - avg_op_ms
    return rec


def bench_sharded(args, dist, rank, world, local_rank, make_layer, workload_config, peaks,
                  ClockSampler, METRIC, UNIT, preflight: Optional[Callable] = None):
    """configs[4]: the 12-mode cutoff-6 state vector, strong scaling over `world` ranks."""
    from . import _lib
    from .core import ffi

    lib = _lib.load()
    want_peers = not getattr(args, "nccl_swap", False)
    parity = preflight("vec", want_peers) if preflight is not None else None
    modes, cutoff = args.modes, args.cutoff
    dims = (cutoff,) * modes
    N = cutoff**modes
    layer = make_layer(modes, cutoff)
    d_ops = [torch.as_tensor(o).cuda() for _, o, _ in layer]
    nops = len(layer)
    vaxes, sparts, shard_a, shard_b = _layouts(ShardedStateVector, dims, world, modes)
    n_local = N // world
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    peers, swap_impl = make_peers(n_local, 2, rank, world, dist, local_rank, want_peers)
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
    runner = LayerRunner(st, layer, d_ops, shard_a, shard_b, _apply_vec)
    total_ms, swap_ms, op_ms, fused_op_ms, launches, clocks, counts = _time_layers(
        args, dist, rank, world, local_rank, runner, st, nops, lib, ClockSampler)
    norm_after = float(st.norm2().item())
    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = 2 * n_local * 16
        # local operators only: the op that carries a fused swap is NVLink-bound and reported apart
        n_local_ops = nops - (1 if counts[2] > 0 else 0)
        avg_op_ms = (op_ms - (fused_op_ms if counts[2] > 0 else 0.0)) / n_local_ops
        achieved = alg_bytes / (avg_op_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": nops * N * args.steps / (total_ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks,
            "e2e": {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "skipped": "end-to-end host-buffer path is measured at N=1 only"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "apply_vec (per-rank shard, local ops of the layer)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_op_ms},
            "cpu_baseline": None,
            "axis_swap": _swap_record(world, n_local, swap_ms, fused_op_ms, avg_op_ms, counts, swap_impl, peers),
            "compute_ms_per_step": op_ms, "norm_after": norm_after, "parity_n": parity,
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
    parity = preflight("rho", want_peers) if preflight is not None else None
    modes, cutoff = args.rho_modes, args.rho_cutoff
    dims = (cutoff,) * modes
    D = cutoff**modes
    layer = make_rho_layer(modes, cutoff)
    d_ops = [torch.as_tensor(o).cuda() for _, o, _ in layer]
    nops = len(layer)
    vaxes, sparts, shard_a, shard_b = _layouts(ShardedDensityMatrix, dims, world, modes)
    n_local = D * D // world
    peers, swap_impl = make_peers(n_local, 2, rank, world, dist, local_rank, want_peers)
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
    runner = LayerRunner(st, layer, d_ops, shard_a, shard_b, _apply_rho)
    total_ms, swap_ms, op_ms, fused_op_ms, launches, clocks, counts = _time_layers(
        args, dist, rank, world, local_rank, runner, st, nops, lib, ClockSampler)
    tr1 = float(st.trace().item())
    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = 2 * n_local * 16
        avg_op_ms = op_ms / nops
        achieved = alg_bytes / (avg_op_ms * 1e-3) / 1e9
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
            "e2e": {"value": None, "unit": RHO_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "skipped": "device-resident workload (the host-buffer path is measured on the headline workload)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "Kraus / U rho U^dagger on a row block (one pass per op)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_op_ms},
            "cpu_baseline": None,
            "axis_swap": _swap_record(world, n_local, swap_ms, fused_op_ms, avg_op_ms, counts, swap_impl, peers),
            "compute_ms_per_step": op_ms, "trace_before": tr0, "trace_after": tr1, "parity_n": parity,
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
