#!/usr/bin/env python
"""
bench.py -- headline benchmark of the photon_weave state-evolution hot path on B200.

Workload (BASELINE.json configs[4], the only config quoted at 1/2/4/8 GPUs and the
largest that fits one GPU): a 12-mode Fock state vector, cutoff 6, complex128
(6^12 = 2.18e9 amplitudes, 34.83 GB).  One STEP is one circuit layer through
``apply_operation``: a single-mode operator on every mode (displacement, phase
shift, squeezing, dense random unitary, cycling) followed by a beam splitter on
every even brick (0,1), (2,3), ... -- 18 operator applications, each updating all
N amplitudes.  metric = amplitudes updated per second (whole job).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "state amplitudes updated/sec (apply_operation layer, 12-mode cutoff-6 state vector)"
UNIT = "amplitudes/s"


# ------------------------------------------------------------------------------------
# workload definition (shared by both arms)
# ------------------------------------------------------------------------------------


def make_layer(modes: int, cutoff: int):
    """[(targets, operator complex128, label)] -- SURVEY.md section 8d operator set."""
    from oracle import operators as ops  # operator *construction* is input generation

    single = [
        ("displace", ops.displacement_operator(cutoff, 0.5)),
        ("phase", ops.phase_operator(cutoff, 0.3)),
        ("squeeze", ops.squeezing_operator(cutoff, 0.2 * (1 + 1j))),
        ("unitary", ops.random_unitary(cutoff, 11)),
    ]
    layer = []
    for m in range(modes):
        name, op = single[m % len(single)]
        layer.append(((m,), np.ascontiguousarray(op), f"{name}[{m}]"))
    bs = np.ascontiguousarray(ops.beam_splitter(cutoff, cutoff, math.pi / 4))
    for m in range(0, modes - 1, 2):
        layer.append(((m, m + 1), bs, f"bs[{m},{m + 1}]"))
    return layer


def make_rho_layer(modes: int, cutoff: int):
    """configs[3] layer: a photon-loss Kraus channel (K = cutoff operators) on every
    mode, then beam splitters on the even bricks.  [(targets, operator(s) complex128, label)]."""
    from oracle import operators as ops  # operator *construction* is input generation

    loss = np.ascontiguousarray(ops.loss_channel(cutoff, 0.1))
    layer = [((m,), loss, f"kraus_loss[{m}]") for m in range(modes)]
    bs = np.ascontiguousarray(ops.beam_splitter(cutoff, cutoff, math.pi / 4))
    for m in range(0, modes - 1, 2):
        layer.append(((m, m + 1), bs, f"bs[{m},{m + 1}]"))
    return layer


def oracle_apply(kind, dims, targets, state, op):
    """The checker of the N>1 parity preflight (oracle/ is test infrastructure: it only CHECKS)."""
    from oracle import adapters_np as oad

    if kind == "vec":
        return oad.apply_operation_vector(tuple(dims), tuple(targets), state, op, True)
    if op.ndim == 3:
        return oad.apply_kraus_matrix(tuple(dims), tuple(targets), state, op, True)
    return oad.apply_operation_matrix(tuple(dims), tuple(targets), state, op, True)


def preflight_input(kind, dims):
    from oracle import operators as ops

    n = int(np.prod(dims))
    return ops.random_state(n, 99) if kind == "vec" else ops.random_density(n, 98)


def ncu_traffic(kernel_key: str):
    """dram__bytes_read+write per launch from the committed `ncu --set full` capture
    (profiles/r2_traffic.json, from tools/ncu_summary.py --by-kernel); None if absent."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        return json.load(open(path)).get(kernel_key, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = f"/tmp/pw_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu), "-lms", "100"],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)),
                       reasons=sorted(reasons), samples=len(sm))
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------------------------
# CPU arm (oracle port), used for cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------


def cpu_layer_throughput(modes: int, cutoff: int, steps: int, warmup: int):
    from oracle import cpu_baseline as cb
    from oracle import operators as ops

    dims = (cutoff,) * modes
    N = cutoff**modes
    layer = [(t, o) for t, o, _ in make_layer(modes, cutoff)]
    psi = ops.random_state(N, 0)
    for _ in range(warmup):
        psi = cb.run_layer(dims, layer, psi)
    t0 = time.perf_counter()
    for _ in range(steps):
        psi = cb.run_layer(dims, layer, psi)
    dt = time.perf_counter() - t0
    return len(layer) * N * steps / dt, dt / steps, len(layer), N


def pick_cpu_modes(cutoff: int, budget_s: float):
    """Largest mode count whose layer is estimated to fit the time budget."""
    from oracle import cpu_baseline as cb
    from oracle import operators as ops

    m = 8
    dims = (cutoff,) * m
    psi = ops.random_state(cutoff**m, 0)
    layer = [(t, o) for t, o, _ in make_layer(m, cutoff)]
    cb.run_layer(dims, layer, psi)
    t0 = time.perf_counter()
    cb.run_layer(dims, layer, psi)
    per_amp_op = (time.perf_counter() - t0) / (len(layer) * cutoff**m)
    best = m
    for cand in (9, 10, 11):
        nops = cand + cand // 2
        if per_amp_op * nops * cutoff**cand <= budget_s:
            best = cand
    return best


def reference_layer_jax(modes, cutoff, steps, warmup):
    """The REAL reference (photon_weave on JAX-CPU) when it is importable on this box:
    adapters.apply_operation_vector_jit (core/adapters.py:752-780), use_contraction=True (the
    library default route), timed like benchmarks/contraction_vs_jitted.py:64-75."""
    import jax  # noqa: F401 -- raises ImportError where jax is absent (this image)
    import jax.numpy as jnp

    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref_dir) and ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    from photon_weave.core import adapters as rad  # the unmodified reference package
    from oracle import operators as ops

    dims = (cutoff,) * modes
    N = cutoff**modes
    layer = [(t, jnp.asarray(o)) for t, o, _ in make_layer(modes, cutoff)]
    psi = jnp.asarray(ops.random_state(N, 0))

    def run(x):
        for t, o in layer:
            x = rad.apply_operation_vector_jit(dims, tuple(t), x, o, True)
        return jax.block_until_ready(x)

    for _ in range(max(warmup, 1)):
        psi = run(psi)
    t0 = time.perf_counter()
    for _ in range(steps):
        psi = run(psi)
    dt = time.perf_counter() - t0
    return len(layer) * N * steps / dt, dt / steps, len(layer), N


def pin_blas_threads():
    """Let the NumPy/BLAS port use every host core, and say how many it got."""
    cores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        threadpool_limits(limits=cores)
        got = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
        return cores, int(got)
    except Exception:
        return cores, None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cutoff = args.cutoff
    cores, blas_threads = pin_blas_threads()
    per_step_budget = 150.0 / max(args.steps + args.warmup, 1)
    modes = pick_cpu_modes(cutoff, per_step_budget)
    kind, what = "port", "NumPy/BLAS port of the reference path (oracle/cpu_baseline.py; jax is not installable here)"
    try:
        value, s_per_step, nops, N = reference_layer_jax(modes, cutoff, args.steps, args.warmup)
        kind, what = "reference", "photon_weave adapters.apply_operation_vector_jit on JAX-CPU (use_contraction=True)"
    except ImportError:
        value, s_per_step, nops, N = cpu_layer_throughput(modes, cutoff, args.steps, args.warmup)
    sample = (f"same layer pattern on a {modes}-mode cutoff-{cutoff} state vector "
              f"({N} amplitudes, {nops} ops/step), {what}")
    # config describes what THIS arm ran: a bounded sample of the workload (per-amplitude
    # throughput is the comparable quantity), not the 12-mode state of the GPU arm
    cfg = {
        "workload": f"bounded CPU sample of configs[4]: {modes}-mode state vector, cutoff {cutoff} "
                    f"({N} amplitudes, {N * 16 / 1e9:.3f} GB complex128)",
        "layer": f"{modes} single-mode ops (displace/phase/squeeze/dense unitary) + "
                 f"{modes // 2} beam splitters on even bricks",
        "ops_per_step": nops, "sample_modes": modes, "sample_amplitudes": N,
        "gpu_arm_modes": args.modes, "same_size_as_gpu_arm": modes == args.modes,
        "comparison": "per-amplitude throughput (amplitudes updated/s); the CPU cannot hold the 34.8 GB state twice",
        "sharding": "host CPU, all cores",
    }
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "blas_threads": blas_threads,
                         "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    N = args.cutoff**args.modes
    return {
        "workload": f"configs[4]: synthetic {args.modes}-mode state vector, cutoff {args.cutoff} "
                    f"({N} amplitudes, {N * 16 / 1e9:.2f} GB complex128)",
        "layer": f"{args.modes} single-mode ops (displace/phase/squeeze/dense unitary) + "
                 f"{args.modes // 2} beam splitters on even bricks",
        "ops_per_step": args.modes + args.modes // 2,
        "sharding": "single device" if world == 1 else f"leading axes block-sharded over {world} ranks",
        "l2": "state (>= 4 GB per rank) exceeds the 126 MB L2; no flush needed",
    }


# ------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------


def run_gpu_arm(args):
    import torch

    import __graft_entry__ as g

    g.build()
    from photon_weave_b200 import _lib
    from photon_weave_b200.core import ffi

    lib = _lib.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world > 1 or args.workload == "rho":
        from photon_weave_b200 import bench_sharded as bs  # noqa: WPS433

        def preflight(kind, want_peers, mode):
            # N>1 only: a small state through the same schedule on the same ranks vs the oracle
            if world == 1 or args.no_preflight:
                return None
            return bs.parity_preflight(dist, rank, world, local_rank, kind,
                                       make_layer if kind == "vec" else make_rho_layer,
                                       oracle_apply, preflight_input, want_peers, mode=mode)

        if args.workload == "rho":
            return bs.bench_sharded_rho(args, dist, rank, world, local_rank, make_rho_layer, peaks,
                                        ClockSampler, preflight)
        return bs.bench_sharded(args, dist, rank, world, local_rank, make_layer, workload_config,
                                peaks, ClockSampler, METRIC, UNIT, preflight)

    modes, cutoff = args.modes, args.cutoff
    dims = (cutoff,) * modes
    N = cutoff**modes
    layer = make_layer(modes, cutoff)
    d_ops = [torch.as_tensor(o).cuda() for _, o, _ in layer]
    targets = [t for t, _, _ in layer]
    nops = len(layer)

    # synthetic state on the device (plumbing), normalised with our reductions
    gen = torch.Generator(device="cuda").manual_seed(0)
    a = torch.view_as_complex(torch.randn(N, 2, generator=gen, device="cuda", dtype=torch.float64))
    stats = ffi.norm_stats(a)
    ffi.scale_(a, (1.0 / torch.sqrt(stats[:1])).contiguous())
    b = torch.empty_like(a)

    def layer_pass(src, dst, events=None):
        for i in range(nops):
            if events is not None:
                events[i].record()
            ffi.apply_vec(src, d_ops[i], dims, targets[i], out=dst)
            src, dst = dst, src
        return src, dst

    for _ in range(args.warmup):
        a, b = layer_pass(a, b)
    torch.cuda.synchronize()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(nops + 1)] for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = lib.pw_launch_count()
    torch.cuda.synchronize()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for s in range(args.steps):
        a, b = layer_pass(a, b, ev[s])
        ev[s][nops].record()
    t_end.record()
    torch.cuda.synchronize()
    launches = int(lib.pw_launch_count() - n0)
    clocks = sampler.stop()
    total_ms = t_start.elapsed_time(t_end)
    ms_per_step = total_ms / args.steps
    value = nops * N * args.steps / (total_ms * 1e-3)

    # norm must still be 1: every operator in the layer is unitary
    norm_after = float(ffi.norm_stats(a)[0])

    # ---- the same layer through the circuit entry point (pw_circuit_device, fuse = 1): commuting
    # single-mode operators are grouped three per HBM pass (resident-tile kernel); results equal the
    # op-by-op layer (tests/test_gpu_parity.py::test_device_circuit_groups_commuting_operators)
    ops_list = [(0, targets[i], d_ops[i]) for i in range(nops)]
    fa, fb = ffi.circuit_device(a, b, False, dims, ops_list, fuse=True)
    torch.cuda.synchronize()
    n_f0 = lib.pw_launch_count()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        fa, fb = ffi.circuit_device(fa, fb, False, dims, ops_list, fuse=True)
    f1.record()
    torch.cuda.synchronize()
    fused_ms = f0.elapsed_time(f1) / args.steps
    fused = {"ms_per_step": fused_ms, "value": nops * N / (fused_ms * 1e-3), "unit": UNIT,
             "launches_per_step": (lib.pw_launch_count() - n_f0) / args.steps,
             "norm_after": float(ffi.norm_stats(fa)[0]),
             "api": "pw_circuit_device(fuse=1): up to 3 commuting single-mode operators per HBM pass "
                    "(pw_apply_vec_multi), beam splitters one pass each"}
    a, b = fa, fb

    # per-operation device times from the events recorded inside the timed region
    per_op_ms = np.zeros(nops)
    for s in range(args.steps):
        for i in range(nops):
            per_op_ms[i] += ev[s][i].elapsed_time(ev[s][i + 1])
    per_op_ms /= args.steps
    single_idx = [i for i in range(nops) if len(targets[i]) == 1]
    pair_idx = [i for i in range(nops) if len(targets[i]) == 2]
    peak, peak_src = peaks()
    alg_bytes = 2 * N * 16
    t_single = float(per_op_ms[single_idx].sum())
    t_pair = float(per_op_ms[pair_idx].sum()) if pair_idx else 0.0
    dom = single_idx if t_single >= t_pair else pair_idx
    dom_name = "apply_vec single-mode (t=%d)" % cutoff if dom is single_idx else \
        "apply_vec two-mode beam splitter (t=%d)" % (cutoff * cutoff)
    dom_ms = float(per_op_ms[dom].mean())
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": ncu_traffic("single_mode" if dom is single_idx else "beam_splitter"),
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": dom_ms,
        "share_of_step": (t_single if dom is single_idx else t_pair) / float(per_op_ms.sum()),
    }
    kernels = {
        "single_mode": {"launches_per_step": len(single_idx),
                        "avg_ms": float(per_op_ms[single_idx].mean()),
                        "min_ms": float(per_op_ms[single_idx].min()),
                        "max_ms": float(per_op_ms[single_idx].max()),
                        "gbs": alg_bytes / (float(per_op_ms[single_idx].mean()) * 1e-3) / 1e9},
    }
    if pair_idx:
        kernels["beam_splitter"] = {
            "launches_per_step": len(pair_idx),
            "avg_ms": float(per_op_ms[pair_idx].mean()),
            "gbs": alg_bytes / (float(per_op_ms[pair_idx].mean()) * 1e-3) / 1e9}
    kernels["per_op_ms"] = {layer[i][2]: round(float(per_op_ms[i]), 4) for i in range(nops)}
    if not args.no_extras and modes >= 5:
        # worst case of the sweep: a dense Haar-random two-mode operator.  FP64-bound
        # (8 t real flop per 32 bytes), served by the FP64 tensor-core kernel (pw_dense.cu).
        from oracle import operators as _ops  # input generation only

        t2 = cutoff * cutoff
        dense2 = torch.as_tensor(np.ascontiguousarray(_ops.random_unitary(t2, 2))).cuda()
        ffi.apply_vec(a, dense2, dims, (3, 4), out=b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ffi.apply_vec(a, dense2, dims, (3, 4), out=b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        kernels["dense_%dx%d_modes_3_4" % (t2, t2)] = {
            "ms": ms, "gbs": alg_bytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": alg_bytes / (ms * 1e-3) / 1e9 / peak,
            "fp64_tflops": 8.0 * N * t2 / (ms * 1e-3) / 1e12,
            "note": "FP64-bound: 8*t real flop per amplitude; mma.sync m8n8k4 f64"}
        del dense2

    # ---- end to end: host buffers through the C ABI (pw_circuit_host) -------------------
    e2e, h_pinned = run_e2e_prepare(args, torch, a, N)
    del a, b
    torch.cuda.empty_cache()
    if h_pinned is not None:
        e2e = run_e2e(args, torch, ffi, layer, h_pinned, N, nops)
        del h_pinned

    # ---- density-matrix side of the metric (configs[3]): Kraus / apply on a 5-mode rho -----
    extras = None if args.no_extras else run_matrix_extras(torch, ffi, peak)
    small = None if args.no_extras else run_small_configs(torch, ffi)

    # ---- CPU baseline on a bounded sample --------------------------------------------------
    cpu = None
    if not args.no_cpu:
        pin_blas_threads()
        m_cpu = pick_cpu_modes(cutoff, 6.0)
        v, s_step, nops_c, n_c = cpu_layer_throughput(m_cpu, cutoff, 2, 1)
        cores, blas_threads = pin_blas_threads()
        cpu = {"value": v, "unit": UNIT, "cores": cores, "blas_threads": blas_threads, "kind": "port",
               "sample": f"same layer pattern on a {m_cpu}-mode cutoff-{cutoff} state vector "
                         f"({n_c} amplitudes), 2 timed steps, NumPy/BLAS port of the reference path"}

    roofline_kraus = None
    if extras:
        # the Kraus side of the metric (configs[3]): loss channel, K = 8, on a middle mode of the
        # 17.18 GB rho -- ONE pass, algorithmic bytes 2 N^2 16 B independent of K
        k = extras["kraus_loss_K8_mode2"]
        roofline_kraus = {
            "bound": "hbm", "kernel": "pw_kraus_mat: loss channel K=8 on mode 2 of the 5-mode cutoff-8 rho "
                                      "(k_build_super + k_apply_blocks, one HBM pass)",
            "achieved": k["gbs"], "peak": peak, "unit": "GB/s", "frac": k["gbs"] / peak,
            "traffic": ncu_traffic("kraus_loss"), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": extras["algorithmic_bytes_per_launch"], "avg_launch_ms": k["ms"],
            "entries_per_s": k["entries_per_s"],
            "innermost_mode": {"ms": extras["kraus_loss_K8_mode4_innermost"]["ms"],
                               "frac": extras["kraus_loss_K8_mode4_innermost"]["frac_of_hbm_peak"]},
        }
    e2e_serial = e2e.get("serial") if isinstance(e2e, dict) else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": workload_config(args, 1), "clocks": clocks, "e2e": e2e, "e2e_serial": e2e_serial,
        "value_fused": fused["value"], "fused": fused,
        "gpu_launches": launches, "roofline": roofline, "roofline_kraus": roofline_kraus,
        "cpu_baseline": cpu,
        "kernels": kernels, "norm_after": norm_after, "density_matrix": extras, "small_configs": small,
        "paths": _lib.path_counters(),
    }
    print(json.dumps(line), flush=True)


def pinned_near_gpu(torch, n):
    """Pinned complex128 host buffer allocated while the process is bound to the CPUs of the
    GPU's NUMA node (NVML affinity; the binding is dropped again right after)."""
    from photon_weave_b200.bench_sharded import _gpu_local_cpus

    before = os.sched_getaffinity(0)
    near = _gpu_local_cpus(torch.cuda.current_device())
    try:
        if near:
            os.sched_setaffinity(0, near)
    except OSError:      # not allowed in this container: allocate wherever the process runs
        near = None
    try:
        return torch.empty(n, dtype=torch.complex128, pin_memory=True)
    finally:
        if near:
            os.sched_setaffinity(0, before)


def run_e2e_prepare(args, torch, dev_state, N):
    """Copy the evolved state into a pinned host buffer (outside any timed region)."""
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    need = N * 16
    if avail < need * 1.5 + (8 << 30):
        return ({"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                 "skipped": f"host has {avail / 1e9:.0f} GB available, needs {need * 1.5 / 1e9:.0f} GB"},
                None)
    h = pinned_near_gpu(torch, N)
    h.copy_(dev_state.reshape(-1))
    torch.cuda.synchronize()
    return None, h


def run_e2e(args, torch, ffi, layer, h, N, nops):
    """Same step, host buffers: H2D(state + operators) + layer + D2H(state) per step.

    Two measurements through the C ABI, both with pinned host buffers and every copy inside
    the timed region:
    * ``serial``: pw_circuit_host, one blocking call per step (latency of one job);
    * ``pipelined`` (the headline ``value``): pw_pipeline_submit / pw_pipeline_wait -- each step
      is still one upload + one layer + one download, but consecutive steps overlap (upload of
      step k+1 while step k computes and step k-1 downloads).  The timed region runs from the
      first submit to the last wait, so pipeline fill and drain are included.
    """
    dims = (args.cutoff,) * args.modes
    need = N * 16
    h_np = h.numpy()
    ops_list = [(0, t, o) for t, o, _ in layer]
    op_bytes = sum(o.nbytes for _, o, _ in layer)
    dev = torch.cuda.current_device()
    steps = max(1, min(args.steps, args.e2e_steps))
    ffi.circuit_host(h_np, False, dims, ops_list, device=dev, out=h_np)  # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        ffi.circuit_host(h_np, False, dims, ops_list, device=dev, out=h_np)
    dt_serial = (time.perf_counter() - t0) / steps
    out = {"value": nops * N / dt_serial, "unit": UNIT, "h2d_bytes_per_step": need + op_bytes,
           "d2h_bytes_per_step": need, "ms_per_step": dt_serial * 1e3, "steps": steps,
           "api": "pw_circuit_host (C ABI, pinned host buffers in and out)"}
    serial = {"ms_per_step": dt_serial * 1e3, "value": nops * N / dt_serial, "steps": steps,
              "api": "pw_circuit_host, one blocking call per step"}

    # streaming pipeline: needs a second pinned buffer (results) and 4 device buffers
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    if args.no_pipeline or avail < need + (16 << 30):
        out["serial"] = serial
        return out
    h_out = pinned_near_gpu(torch, N)
    o_np = h_out.numpy()
    psteps = max(steps, args.e2e_pipeline_steps)
    with ffi.HostPipeline(dims, False, np.complex128, device=dev, buffers=4) as pl:
        pl.wait(pl.submit(h_np, o_np, ops_list))  # warm-up
        t0 = time.perf_counter()
        pending = []
        for _ in range(psteps):
            pending.append(pl.submit(h_np, o_np, ops_list))
            if len(pending) > 2:  # at most three steps in flight
                pl.wait(pending.pop(0))
        for tk in pending:
            pl.wait(tk)
        dt = (time.perf_counter() - t0) / psteps
        nbuf = pl.buffers
    check = float(np.vdot(o_np[:1 << 20], o_np[:1 << 20]).real)
    del h_out
    out.update({"value": nops * N / dt, "ms_per_step": dt * 1e3, "steps": psteps,
                "api": "pw_pipeline_submit/wait (C ABI, pinned host buffers): THROUGHPUT over INDEPENDENT "
                       "JOBS -- each step is a separate job (its own upload, layer and download; the same "
                       "input buffer re-submitted), up to three in flight, so the upload of job k+1, the "
                       "kernels of job k and the download of job k-1 overlap; fill and drain timed.  "
                       "The latency of ONE job is e2e_serial (pw_circuit_host).",
                "mode": "pipelined, independent jobs",
                "device_buffers": nbuf, "serial": serial, "checksum_first_1Mi": check})
    return out


def run_matrix_extras(torch, ffi, peak):
    """configs[3]: 5-mode cutoff-8 density matrix (17.18 GB c128): Kraus loss channel and
    operator applications, device-resident, CUDA-event timed (informational)."""
    from oracle import operators as ops

    d, m = 8, 5
    dims = (d,) * m
    D = d**m
    gen = torch.Generator(device="cuda").manual_seed(1)
    rho = torch.view_as_complex(torch.randn(D * D, 2, generator=gen, device="cuda", dtype=torch.float64))
    out = torch.empty_like(rho)
    nbytes = 2 * D * D * 16

    def dev(x):
        return torch.as_tensor(np.ascontiguousarray(x)).cuda()

    loss = dev(ops.loss_channel(d, 0.1))
    u = dev(ops.random_unitary(d, 1))
    ph = dev(ops.phase_operator(d, 0.3))
    bs = dev(ops.beam_splitter(d, d, math.pi / 4))

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    res = {"workload": "configs[3]: 5-mode cutoff-8 density matrix (2^30 entries, 17.18 GB complex128)",
           "algorithmic_bytes_per_launch": nbytes}
    res["kraus_loss_K8_mode2"] = None
    ms = timed(lambda: ffi.kraus_mat(rho, loss, dims, (2,), out=out), reps=8)
    gbs = nbytes / (ms * 1e-3) / 1e9
    res["kraus_loss_K8_mode2"] = {"ms": ms, "entries_per_s": D * D / (ms * 1e-3), "gbs": gbs,
                                  "frac_of_hbm_peak": gbs / peak, "reps": 8}
    for name, fn in [
        ("apply_phase_mode1", lambda: ffi.apply_mat(rho, ph, dims, (1,), out=out)),
        ("apply_dense_8x8_mode2", lambda: ffi.apply_mat(rho, u, dims, (2,), out=out)),
        ("apply_dense_8x8_mode4_innermost", lambda: ffi.apply_mat(rho, u, dims, (4,), out=out)),
        ("apply_beam_splitter_modes_1_2", lambda: ffi.apply_mat(rho, bs, dims, (1, 2), out=out)),
        ("apply_beam_splitter_modes_3_4_innermost", lambda: ffi.apply_mat(rho, bs, dims, (3, 4), out=out)),
        ("kraus_loss_K8_mode4_innermost", lambda: ffi.kraus_mat(rho, loss, dims, (4,), out=out)),
    ]:
        ms = timed(fn)
        gbs = nbytes / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "entries_per_s": D * D / (ms * 1e-3), "gbs": gbs, "frac_of_hbm_peak": gbs / peak}
    ms = timed(lambda: ffi.probs(rho.reshape(D, D), dims, (1,), True), reps=5)
    res["probs_mat_mode1"] = {"ms": ms}
    ms = timed(lambda: ffi.trace_out_mat(rho.reshape(D, D), dims, (1,)), reps=5)
    res["trace_out_keep_mode1"] = {"ms": ms}
    del rho, out
    torch.cuda.empty_cache()
    return res


def run_small_configs(torch, ffi):
    """configs[0..2] (launch-latency bound, L2 resident): microseconds per operation through
    the seam, next to the NumPy port on the host (informational)."""
    import math as _m

    from oracle import adapters_np as oad
    from oracle import operators as ops
    from photon_weave_b200.core import adapters as ad

    def dev(x):
        # operators of a circuit are immutable: pinned, so the library keeps their structure analysis
        # (what the container layer does for every cached Operation.device_operator)
        t = torch.as_tensor(np.ascontiguousarray(x)).cuda()
        return ffi.pin_operator(t) if t.dim() >= 2 and t.shape[-1] == t.shape[-2] and t.shape[-1] <= 128 else t

    def gpu_us(fn, reps=200):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e6

    def cpu_us(fn, reps=50):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e6

    res = {}
    # configs[0] Mach-Zehnder: 2 Fock modes (cutoff 2), BS - phase - BS, then two reduced states
    dims = (2, 2)
    bs, ph = ops.beam_splitter(2, 2, _m.pi / 4), ops.phase_operator(2, 0.7)
    psi = np.zeros((4, 1), dtype=np.complex128); psi[2, 0] = 1
    d_psi, d_bs, d_ph = dev(psi), dev(bs), dev(ph)

    def mzi_gpu():
        x = ad.apply_operation_vector(dims, (0, 1), d_psi, d_bs)
        x = ad.apply_operation_vector(dims, (0,), x, d_ph)
        x = ad.apply_operation_vector(dims, (0, 1), x, d_bs)
        return ad.trace_out_vector(dims, (0,), x), ad.trace_out_vector(dims, (1,), x)

    def mzi_cpu():
        x = oad.apply_operation_vector(dims, (0, 1), psi, bs, True)
        x = oad.apply_operation_vector(dims, (0,), x, ph, True)
        x = oad.apply_operation_vector(dims, (0, 1), x, bs, True)
        return oad.trace_out_vector(dims, (0,), x, True), oad.trace_out_vector(dims, (1,), x, True)

    res["mach_zehnder_5_calls"] = {"gpu_us": gpu_us(mzi_gpu), "cpu_port_us": cpu_us(mzi_cpu)}
    # the same five calls captured once and replayed as ONE CUDA graph launch (core/graphs.py)
    from photon_weave_b200.core.graphs import GraphedCalls

    def mzi_seq(x):
        x = ad.apply_operation_vector(dims, (0, 1), x, d_bs)
        x = ad.apply_operation_vector(dims, (0,), x, d_ph)
        x = ad.apply_operation_vector(dims, (0, 1), x, d_bs)
        return ad.trace_out_vector(dims, (0,), x), ad.trace_out_vector(dims, (1,), x)

    mzi_graph = GraphedCalls(mzi_seq, d_psi)
    res["mach_zehnder_5_calls"]["gpu_graph_replay_us"] = gpu_us(lambda: mzi_graph())
    # configs[1] lossy 4-mode circuit on a density matrix, cutoff 5 (6.25 MB): BS + loss channels
    dims = (5, 5, 5, 5)
    D = 625
    rho = ops.random_density(D, 3)
    bs5, loss = ops.beam_splitter(5, 5, _m.pi / 4), ops.loss_channel(5, 0.1)
    d_rho, d_bs5, d_loss = dev(rho), dev(bs5), dev(loss)
    res["lossy_4mode_rho_bs"] = {
        "gpu_us": gpu_us(lambda: ad.apply_operation_matrix(dims, (1, 2), d_rho, d_bs5), 50),
        "cpu_port_us": cpu_us(lambda: oad.apply_operation_matrix(dims, (1, 2), rho, bs5, True), 3)}
    res["lossy_4mode_rho_kraus_K5"] = {
        "gpu_us": gpu_us(lambda: ad.apply_kraus_matrix(dims, (2,), d_rho, d_loss), 50),
        "cpu_port_us": cpu_us(lambda: oad.apply_kraus_matrix(dims, (2,), rho, loss, True), 3)}
    # configs[2] Jaynes-Cummings: cutoff 64 (x) qubit, 10k steps of (apply, trace_out(qubit))
    dims = (64, 2)
    U = ops.random_unitary(128, 5)
    psi = ops.random_state(128, 6)
    d_U, d_psi = dev(U), dev(psi).reshape(-1)
    nsteps = 10000
    ffi.evolve_vec(d_psi, d_U, dims, 10, keep=(1,))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, red = ffi.evolve_vec(d_psi, d_U, dims, nsteps, keep=(1,))
    red_host = red.cpu()
    fused_s = time.perf_counter() - t0

    def jc_step_gpu():
        x = ad.apply_operation_vector(dims, (0, 1), d_psi.reshape(-1, 1), d_U)
        return ad.trace_out_vector(dims, (1,), x)

    def jc_step_cpu():
        x = U @ psi
        return oad.trace_out_vector(dims, (1,), x, True)

    # a BATCH of trajectories under the same propagator (vmap): one dense tensor-core launch per step
    from photon_weave_b200.core import transforms as T

    Bt = 1 << 18
    gen = torch.Generator(device="cuda").manual_seed(9)
    batch = torch.view_as_complex(torch.randn(Bt, 128, 1, 2, generator=gen, device="cuda", dtype=torch.float64))
    step = T.vmap(lambda s, o: ad.apply_operation_vector(dims, (0, 1), s, o), in_axes=(0, None))
    step(batch, d_U)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        batch = step(batch, d_U)
    e1.record()
    torch.cuda.synchronize()
    ms_b = e0.elapsed_time(e1) / 5
    res["jaynes_cummings_batched"] = {
        "trajectories": Bt, "ms_per_step": ms_b, "trajectory_steps_per_s": Bt / (ms_b * 1e-3),
        "fp64_tflops": 8.0 * 128 * 128 * Bt / (ms_b * 1e-3) / 1e12,
        "note": "128x128 propagator on 2^18 states = GEMM on the FP64 tensor cores (k_apply_dense_mma, t = 128); "
                "measured DMMA roof 37.1 TFLOP/s (profiles/r2_arith_peaks.json)"}
    del batch
    jc_graph = GraphedCalls(lambda p: (lambda x: (x, ad.trace_out_vector(dims, (1,), x)))(
        ad.apply_operation_vector(dims, (0, 1), p, d_U)), d_psi.reshape(-1, 1).clone())
    res["jaynes_cummings_10k_steps"] = {
        "fused_kernel_total_ms": fused_s * 1e3, "fused_us_per_step": fused_s / nsteps * 1e6,
        "per_call_api_us_per_step": gpu_us(jc_step_gpu, 200),
        "graph_replay_us_per_step": gpu_us(lambda: jc_graph(), 200),
        "cpu_port_us_per_step": cpu_us(jc_step_cpu, 200),
        "norm_check": float(np.trace(red_host[-1].numpy()).real)}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--modes", type=int, default=12)
    ap.add_argument("--cutoff", type=int, default=6)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-pipeline-steps", type=int, default=16,
                    help="steps (independent jobs) of the streaming end-to-end measurement; pipeline fill "
                         "and drain (one upload + one download that overlap nothing) are inside the timed "
                         "region, so fewer steps quote more of them per step")
    ap.add_argument("--no-pipeline", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--nccl-swap", action="store_true", help="N>1: NCCL all_to_all instead of peer reads")
    ap.add_argument("--workload", default="vec", choices=["vec", "rho"],
                    help="vec: configs[4] 12-mode state vector (headline); rho: configs[3] 5-mode "
                         "density matrix, Kraus + beam-splitter layer (row-sharded for N>1)")
    ap.add_argument("--rho-modes", type=int, default=5)
    ap.add_argument("--rho-cutoff", type=int, default=8)
    ap.add_argument("--no-preflight", action="store_true", help="N>1: skip the parity preflight")
    ap.add_argument("--swap", default="pipelined", choices=["pipelined", "fused", "plain"],
                    help="N>1 layer schedule: swap hidden behind chunk-wise local operations (default); "
                         "fused gather+apply kernel; stand-alone swap")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
