"""
Multi-GPU state vectors: one process per GPU, the state block-sharded over its
leading subsystem axes, operators on local axes applied without communication, and
an all-to-all AXIS SWAP (NCCL over NVLink) when an operator touches a sharded axis.

The reference has no distributed code at all; this is the
B200-native design of BASELINE.json's north_star.

Layout bookkeeping
------------------
The local buffer is a row-major tensor over *virtual axes*.  A virtual axis
``VAxis(g, extent, weight)`` carries part of global subsystem ``g``'s index:

    index_g = sum over local virtual axes of g (coordinate * weight)
            + sum over sharded parts of g     (rank digit  * weight)

``sparts`` lists the parts that are sharded (their coordinate is a digit of the
rank, most significant first).  An operator on subsystem ``g`` is applied to all
local virtual axes of ``g`` at once (ordered by weight, descending), which is exactly
the multi-target strided kernel of libpwb200 -- no local transpose is ever needed.

Axis swap (``relayout``): pick virtual axes whose extents multiply to the world
size, bring them to the front of the buffer (a pack pass, skipped when they already
lead -- the steady state when two layouts alternate), run ONE all_to_all_single over
all ranks, and relabel: the received leading axes are the previously sharded parts
(now local), the picked axes become the rank digits.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, List, Optional, Sequence, Tuple

import torch


@dataclass(frozen=True)
class VAxis:
    g: int        # global subsystem index
    extent: int
    weight: int   # multiplier of this coordinate inside subsystem g's index


class _RawCuda:
    """Zero-copy torch view of a raw device pointer (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, numel: int, typestr: str):
        self.__cuda_array_interface__ = {
            "shape": (numel,), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None,
        }


class PeerBuffers:
    """
    Ping-pong shard buffers allocated by libpwb200 (cudaMalloc) and exported with CUDA IPC,
    so every rank can read every other rank's shard directly over NVLink.
    """

    def __init__(self, numel: int, dtype: torch.dtype, nbuf: int, rank: int, world: int, dist: Any,
                 device: int):
        import ctypes as C

        from . import _lib

        lib = _lib.load()
        self.rank, self.world, self.numel, self.dtype = rank, world, numel, dtype
        self.esize = 16 if dtype == torch.complex128 else 8
        self._lib = lib
        self.local_ptrs: List[int] = []
        handles = []
        for _ in range(nbuf):
            ptr = C.c_void_p()
            h = (C.c_ubyte * 64)()
            _lib.check(lib.pw_peer_alloc(numel * self.esize, C.byref(ptr), h), "pw_peer_alloc")
            self.local_ptrs.append(int(ptr.value))
            handles.append(bytes(h))
        gathered: List[Any] = [None] * world
        dist.all_gather_object(gathered, handles)
        self.ptrs: List[List[int]] = []  # [buffer][rank]
        self._opened: List[int] = []
        for b in range(nbuf):
            row = []
            for r in range(world):
                if r == rank:
                    row.append(self.local_ptrs[b])
                else:
                    ptr = C.c_void_p()
                    h = (C.c_ubyte * 64).from_buffer_copy(gathered[r][b])
                    _lib.check(lib.pw_peer_open(h, C.byref(ptr)), "pw_peer_open")
                    row.append(int(ptr.value))
                    self._opened.append(int(ptr.value))
            self.ptrs.append(row)
        typestr = "<c16" if dtype == torch.complex128 else "<c8"
        self.tensors = [
            torch.as_tensor(_RawCuda(p, numel, typestr), device=torch.device("cuda", device))
            for p in self.local_ptrs
        ]

    def index_of(self, t: torch.Tensor) -> int:
        return self.local_ptrs.index(t.data_ptr())

    def chunk_ptrs(self, buf: int, chunk_elems: int) -> List[int]:
        """Pointer to chunk #rank inside every rank's buffer `buf`."""
        off = self.rank * chunk_elems * self.esize
        return [p + off for p in self.ptrs[buf]]

    def close(self) -> None:
        for p in self._opened:
            self._lib.pw_peer_close(p)
        self._opened = []
        for p in self.local_ptrs:
            self._lib.pw_peer_free(p)
        self.local_ptrs = []


class CudaBackend:
    """Local arithmetic through libpwb200 (the product path)."""

    def gather_chunks(self, srcs: List[int], chunk_elems: int, out: torch.Tensor, rank: int = 0) -> None:
        import ctypes as C

        from . import _lib
        from .core import _arrays as A

        lib = _lib.load()
        arr = (C.c_void_p * len(srcs))(*srcs)
        _lib.check(lib.pw_gather_chunks(arr, len(srcs), chunk_elems, rank, out.data_ptr(), A.code(out.dtype),
                                        A.stream_ptr()), "pw_gather_chunks")

    def apply_vec_gather(self, srcs: List[int], chunk_elems: int, out: torch.Tensor, op: torch.Tensor,
                         dims, targets, rank: int = 0) -> bool:
        """Fused swap + apply; False when the shape is outside the fused kernel (t > 8)."""
        import ctypes as C

        from . import _lib
        from .core import _arrays as A

        lib = _lib.load()
        arr = (C.c_void_p * len(srcs))(*srcs)
        rc = lib.pw_apply_vec_gather(arr, len(srcs), chunk_elems, rank, out.data_ptr(), op.data_ptr(),
                                     _lib.i64_array(dims), len(dims), _lib.i32_array(targets),
                                     len(targets), A.code(out.dtype), A.stream_ptr())
        if rc == 2:  # PW_ERR_UNSUPPORTED
            return False
        _lib.check(rc, "pw_apply_vec_gather")
        return True

    def apply_vec(self, local: torch.Tensor, op: torch.Tensor, dims, targets, out=None):
        from .core import ffi

        return ffi.apply_vec(local, op, dims, targets, out=out)

    def permute(self, local: torch.Tensor, dims, perm, out=None):
        from .core import ffi

        return ffi.permute(local, dims, perm, out=out)

    def norm2(self, local: torch.Tensor) -> torch.Tensor:
        from .core import ffi

        return ffi.norm_stats(local)[:1].clone()


def rank_digits(rank: int, extents: Sequence[int]) -> List[int]:
    digits = []
    for e in reversed(extents):
        digits.append(rank % e)
        rank //= e
    return list(reversed(digits))


class ShardedStateVector:
    """A state vector of ``global_dims`` block-sharded over ``world`` ranks."""

    def __init__(self, global_dims: Sequence[int], local: torch.Tensor, vaxes: List[VAxis],
                 sparts: List[VAxis], rank: int, world: int, dist: Any, backend: Any,
                 spare: Optional[torch.Tensor] = None, peers: Optional[PeerBuffers] = None):
        self.global_dims = tuple(int(d) for d in global_dims)
        self.local = local.reshape(-1)
        self.vaxes = list(vaxes)
        self.sparts = list(sparts)
        self.rank, self.world = rank, world
        self.dist = dist
        self.backend = backend
        self.spare = spare  # second buffer for out-of-place apply / all-to-all
        self.peers = peers  # IPC-mapped shard buffers: swap = direct NVLink reads, no NCCL
        self.swaps = 0
        self.packs = 0
        self.fused_swaps = 0
        self._flag = None
        assert math.prod(s.extent for s in self.sparts) == world
        assert math.prod(v.extent for v in self.vaxes) == self.local.numel()

    # ---- construction -----------------------------------------------------------------

    @classmethod
    def shard_plan(cls, global_dims: Sequence[int], world: int) -> List[Tuple[int, int]]:
        """[(axis, factor)]: factor the world size over the leading axes (6 = 2 x 3 ...)."""
        plan, rem = [], world
        for axis, d in enumerate(global_dims):
            if rem == 1:
                break
            f = math.gcd(d, rem)
            if f > 1:
                plan.append((axis, f))
                rem //= f
        if rem != 1:
            raise ValueError(f"cannot factor world size {world} over dims {tuple(global_dims)}")
        return plan

    @classmethod
    def initial_layout(cls, global_dims: Sequence[int], world: int):
        plan = dict(cls.shard_plan(global_dims, world))
        vaxes, sparts = [], []
        for g, d in enumerate(global_dims):
            f = plan.get(g, 1)
            if f > 1:
                sparts.append(VAxis(g, f, d // f))
            vaxes.append(VAxis(g, d // f, 1))
        return vaxes, sparts

    @classmethod
    def from_full(cls, global_dims, full: torch.Tensor, rank: int, world: int, dist, backend):
        """Every rank holds the full vector (tests / small cases): keep only the own block."""
        vaxes, sparts = cls.initial_layout(global_dims, world)
        t = full.reshape(tuple(global_dims))
        digits = rank_digits(rank, [s.extent for s in sparts])
        index = [slice(None)] * len(global_dims)
        for s, b in zip(sparts, digits):
            blk = global_dims[s.g] // s.extent
            index[s.g] = slice(b * blk, (b + 1) * blk)
        local = t[tuple(index)].clone().reshape(-1)  # never alias the caller's buffer
        return cls(global_dims, local, vaxes, sparts, rank, world, dist, backend)

    # ---- helpers ----------------------------------------------------------------------

    @property
    def local_dims(self) -> Tuple[int, ...]:
        return tuple(v.extent for v in self.vaxes)

    def sharded_axes(self) -> List[int]:
        return sorted({s.g for s in self.sparts})

    def _targets_for(self, axes: Sequence[int]) -> List[int]:
        targets: List[int] = []
        for g in axes:
            parts = [(v.weight, i) for i, v in enumerate(self.vaxes) if v.g == g]
            parts.sort(reverse=True)
            # the local parts must tile subsystem g's index as a mixed radix number
            expect = 1
            for w, i in sorted(parts):
                if w != expect:
                    raise RuntimeError(f"subsystem {g} is not fully local (layout {self.vaxes})")
                expect *= self.vaxes[i].extent
            if expect != self.global_dims[g]:
                raise RuntimeError(f"subsystem {g} is not fully local")
            targets += [i for _, i in parts]
        return targets

    def _buffers(self) -> torch.Tensor:
        if self.spare is None or self.spare.numel() != self.local.numel():
            self.spare = torch.empty_like(self.local)
        return self.spare

    # ---- operations -------------------------------------------------------------------

    def _barrier(self) -> None:
        """Stream-ordered cross-rank barrier (a one-element all-reduce; no host sync)."""
        if self._flag is None:
            self._flag = torch.zeros(1, dtype=torch.float32, device=self.local.device)
        self.dist.all_reduce(self._flag)

    def apply_operation(self, operator: torch.Tensor, axes: Sequence[int],
                        prefer_shard: Optional[Sequence[int]] = None) -> None:
        """psi <- (operator on global subsystems ``axes``) psi; swaps axes in if needed."""
        if self.peers is not None and any(s.g in axes for s in self.sparts):
            t = math.prod(self.global_dims[a] for a in axes)
            if t <= 8:
                # fused: the operator is applied while the swapped data streams in over NVLink
                self.relayout(self._choose_parts(set(axes), prefer_shard), fused=(operator, axes))
                return
        self.ensure_local(axes, prefer_shard)
        targets = self._targets_for(axes)
        out = self._buffers()
        self.backend.apply_vec(self.local, operator, self.local_dims, targets, out=out)
        self.local, self.spare = out, self.local

    def ensure_local(self, axes: Sequence[int], prefer_shard: Optional[Sequence[int]] = None) -> None:
        if any(s.g in axes for s in self.sparts):
            self.relayout(self._choose_parts(set(axes), prefer_shard))

    def _choose_parts(self, needed: set, prefer: Optional[Sequence[int]]) -> List[int]:
        """
        Choose virtual axes (splitting when necessary) whose extents multiply to the
        world size; returns their positions.  Candidates: axes of the preferred
        subsystems first, then front to back; never a needed subsystem.
        """
        order = list(range(len(self.vaxes)))
        if prefer:
            pref = [i for i in order if self.vaxes[i].g in prefer]
            order = pref + [i for i in order if i not in pref]
        rem = self.world
        chosen: List[int] = []
        splits: List[Tuple[int, int]] = []  # (position, factor to split off the top)
        for i in order:
            if rem == 1:
                break
            v = self.vaxes[i]
            if v.g in needed:
                continue
            f = math.gcd(v.extent, rem)
            if f == 1:
                continue
            if f == v.extent:
                chosen.append(i)
            else:
                splits.append((i, f))
            rem //= f
        if rem != 1:
            raise RuntimeError("no admissible axis swap for this operator / world size")
        # materialise splits (block split: (e) -> (f, e/f), weights (w*e/f, w)); positions shift
        for i, f in sorted(splits, reverse=True):
            v = self.vaxes[i]
            hi = VAxis(v.g, f, v.weight * (v.extent // f))
            lo = VAxis(v.g, v.extent // f, v.weight)
            self.vaxes[i:i + 1] = [hi, lo]
            chosen = [c + 1 if c > i else c for c in chosen]
            chosen.append(i)
        return sorted(chosen)

    def relayout(self, positions: Sequence[int], fused=None) -> None:
        """All-to-all axis swap: the virtual axes at ``positions`` become the rank digits.
        ``fused=(operator, axes)``: apply that operator while gathering (peer mode only)."""
        positions = list(positions)
        k = len(positions)
        assert math.prod(self.vaxes[p].extent for p in positions) == self.world
        if positions != list(range(k)):
            # pack: bring the outgoing parts to the front (one local pass)
            perm = positions + [i for i in range(len(self.vaxes)) if i not in positions]
            out = self._buffers()
            self.backend.permute(self.local, self.local_dims, perm, out=out)
            self.local, self.spare = out, self.local
            self.vaxes = [self.vaxes[i] for i in perm]
            self.packs += 1
        recv = self._buffers()
        new_sparts = self.vaxes[:k]
        new_vaxes = [VAxis(s.g, s.extent, s.weight) for s in self.sparts] + self.vaxes[k:]
        if self.peers is not None:
            chunk = self.local.numel() // self.world
            self._barrier()  # every rank has finished writing the buffer we are about to read
            srcs = self.peers.chunk_ptrs(self.peers.index_of(self.local), chunk)
            done = False
            if fused is not None:
                op, axes = fused
                self.vaxes, self.sparts = new_vaxes, new_sparts
                done = self.backend.apply_vec_gather(srcs, chunk, recv, op, self.local_dims,
                                                     self._targets_for(axes), rank=self.rank)
                if done:
                    self.fused_swaps += 1
            if not done:
                self.backend.gather_chunks(srcs, chunk, recv, rank=self.rank)
            self._barrier()  # peers are done reading before anyone overwrites the old buffer
            self.local, self.spare = recv, self.local
            self.vaxes, self.sparts = new_vaxes, new_sparts
            self.swaps += 1
            if fused is not None and not done:
                self.apply_operation(fused[0], fused[1])
            return
        # complex buffers travel as (n, 2) reals: equal row splits, one piece per rank
        self.dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(self.local))
        self.local, self.spare = recv, self.local
        self.vaxes, self.sparts = new_vaxes, new_sparts
        self.swaps += 1
        if fused is not None:
            self.apply_operation(fused[0], fused[1])

    def norm2(self) -> torch.Tensor:
        """sum |psi|^2 over all ranks (device tensor)."""
        n = self.backend.norm2(self.local)
        self.dist.all_reduce(n)
        return n

    def gather_full(self) -> torch.Tensor:
        """Reassemble the full vector in canonical subsystem order on every rank (tests)."""
        pieces = [torch.empty_like(self.local) for _ in range(self.world)]
        self.dist.all_gather([torch.view_as_real(p) for p in pieces],
                             torch.view_as_real(self.local.contiguous()))
        n = len(self.global_dims)
        full = torch.empty(self.global_dims, dtype=self.local.dtype, device=self.local.device)
        sp_ext = [s.extent for s in self.sparts]
        for r, piece in enumerate(pieces):
            digits = rank_digits(r, sp_ext)
            t = piece.reshape(self.local_dims)
            # index of subsystem g = sum(coord * weight): scatter through an index grid
            idx = [torch.zeros((), dtype=torch.long, device=t.device) for _ in range(n)]
            for s, b in zip(self.sparts, digits):
                idx[s.g] = idx[s.g] + b * s.weight
            nd = len(self.vaxes)
            for pos, v in enumerate(self.vaxes):
                shape = [1] * nd
                shape[pos] = v.extent
                coord = torch.arange(v.extent, device=t.device).reshape(shape) * v.weight
                idx[v.g] = idx[v.g] + coord
            idx = [torch.broadcast_to(i, t.shape) for i in idx]
            full[tuple(idx)] = t
        return full.reshape(-1)


# --------------------------------------------------------------------------------------
# bench.py --gpus N driver (strong scaling of the 12-mode layer)
# --------------------------------------------------------------------------------------


def bench_sharded(args, dist, rank, world, local_rank, make_layer, workload_config, peaks,
                  ClockSampler, METRIC, UNIT):
    import json

    from . import _lib
    from .core import ffi

    lib = _lib.load()
    modes, cutoff = args.modes, args.cutoff
    dims = (cutoff,) * modes
    N = cutoff**modes
    layer = make_layer(modes, cutoff)
    d_ops = [torch.as_tensor(o).cuda() for _, o, _ in layer]
    nops = len(layer)

    vaxes, sparts = ShardedStateVector.initial_layout(dims, world)
    n_local = N // world
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    peers = None
    swap_impl = "torch.distributed.all_to_all_single (NCCL)"
    if not getattr(args, "nccl_swap", False):
        try:
            peers = PeerBuffers(n_local, torch.complex128, 2, rank, world, dist, local_rank)
            swap_impl = ("libpwb200 peer reads over NVLink (CUDA IPC): pw_apply_vec_gather fused "
                         "swap+apply, pw_gather_chunks otherwise")
        except Exception as exc:  # IPC unavailable: fall back to the NCCL collective
            peers = None
            swap_impl += f" [peer memory unavailable: {exc}]"
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

    # Two layouts alternate: A shards the leading axes, B shards axes in the middle, so
    # every brick of the layer is local in one of them and each layer costs ONE swap.
    k = len(st.sparts)
    shard_a = [s.g for s in st.sparts]
    mid = modes // 2
    shard_b = list(range(mid, mid + k))
    ops_a = [i for i, (t, _, _) in enumerate(layer) if not set(t) & set(shard_a)]
    ops_b = [i for i in range(nops) if i not in ops_a]
    assert all(not set(layer[i][0]) & set(shard_b) for i in ops_b)

    def run_layer(events=None):
        # phase 1: everything local under the current layout; then swap; then the rest
        cur_sharded = set(st.sharded_axes())
        first = [i for i in range(nops) if not set(layer[i][0]) & cur_sharded]
        second = [i for i in range(nops) if i not in first]
        other = shard_b if cur_sharded == set(shard_a) else shard_a
        slot = 0
        for i in first:
            if events is not None:
                events[slot].record(); slot += 1
            st.apply_operation(d_ops[i], layer[i][0])
        if events is not None:
            events[slot].record(); slot += 1
        if second and st.peers is None:
            st.ensure_local(layer[second[0]][0] + tuple(a for i in second for a in layer[i][0]),
                            prefer_shard=other)
        for i in second:
            if events is not None:
                events[slot].record(); slot += 1
            # peer mode: the first of these ops carries the swap (fused gather + apply)
            st.apply_operation(d_ops[i], layer[i][0], prefer_shard=other)
        if events is not None:
            events[slot].record()
        return len(first)

    for _ in range(args.warmup):
        run_layer()
    torch.cuda.synchronize()
    dist.barrier()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(nops + 2)] for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = lib.pw_launch_count()
    swaps0, packs0, fused0 = st.swaps, st.packs, st.fused_swaps
    torch.cuda.synchronize()
    dist.barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    nfirst = []
    for s in range(args.steps):
        nfirst.append(run_layer(ev[s]))
    t1.record()
    torch.cuda.synchronize()
    dist.barrier()
    launches = int(lib.pw_launch_count() - n0)
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([t0.elapsed_time(t1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())

    # swap time = the slot between the two phases
    swap_ms = 0.0
    op_ms = 0.0
    fused_op_ms = 0.0
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
    fused_op_ms /= args.steps
    swap_ms /= args.steps
    op_ms /= args.steps
    tm = torch.tensor([swap_ms, op_ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    swap_ms, op_ms = float(tm[0]), float(tm[1])
    norm_after = float(st.norm2().item())
    swaps = (st.swaps - swaps0) / args.steps
    packs = (st.packs - packs0) / args.steps

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = 2 * n_local * 16
        avg_op_ms = op_ms / nops
        achieved = alg_bytes / (avg_op_ms * 1e-3) / 1e9
        swap_bytes = n_local * 16 * (world - 1) / world
        fused_per_step = (st.fused_swaps - fused0) / args.steps
        line = {
            "metric": METRIC, "value": nops * N * args.steps / (total_ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks,
            "e2e": {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "skipped": "end-to-end host-buffer path is measured at N=1 only"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "apply_vec (per-rank shard, all ops of the layer)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_op_ms},
            "cpu_baseline": None,
            "axis_swap": {"swaps_per_step": swaps, "packs_per_step": packs, "ms_per_step": swap_ms,
                          "bytes_sent_per_rank": swap_bytes,
                          # stand-alone swaps only: a swap fused into the apply kernel has no
                          # duration of its own (fused_swap_plus_apply_ms is the fused kernel)
                          "gbs_per_rank_per_direction": swap_bytes * (swaps - fused_per_step) / (swap_ms * 1e-3) / 1e9
                          if swap_ms > 0 and swaps - fused_per_step > 0 else None,
                          "nvlink_reference_gbs": 770.0,
                          "fused_swaps_per_step": fused_per_step,
                          "fused_swap_plus_apply_ms": fused_op_ms if peers is not None else None,
                          "implementation": swap_impl},
            "compute_ms_per_step": op_ms, "norm_after": norm_after,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    if peers is not None:
        del st, local, spare
        torch.cuda.synchronize()
        peers.close()
    dist.destroy_process_group()
