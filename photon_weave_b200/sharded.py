"""
Multi-GPU states: one process per GPU, the state block-sharded over its leading subsystem
axes, operators on local axes applied without communication, and an all-to-all AXIS SWAP
(peer reads over NVLink, or NCCL) when an operator touches a sharded axis.

The reference has no distributed code at all; this is the B200-native
design of BASELINE.json's north_star.  Per block the arithmetic is the
reference's: apply ``core/kernels.py:62-98``, U rho U^dagger / Kraus ``:101-186``, measurement
probabilities ``:28-59``, collapse ``:189-227``, partial trace ``:301-319``.

Layout bookkeeping
------------------
The local buffer is a row-major tensor over *virtual axes*.  A virtual axis
``VAxis(g, extent, weight)`` carries part of global subsystem ``g``'s index:

    index_g = sum over local virtual axes of g (coordinate * weight)
            + sum over sharded parts of g     (rank digit  * weight)

``sparts`` lists the parts that are sharded (their coordinate is a digit of the rank, most
significant first).  An operator on subsystem ``g`` is applied to all local virtual axes of
``g`` at once (ordered by weight, descending), which is exactly the multi-target strided
kernel of libpwb200 -- no local transpose is ever needed.

Axis swap (``relayout``): pick virtual axes whose extents multiply to the world size, bring
them to the front of the buffer (a pack pass, skipped when they already lead -- the steady
state when two layouts alternate), run ONE all-to-all over all ranks, and relabel: the
received leading axes are the previously sharded parts (now local), the picked axes become
the rank digits.

A density matrix is the same machinery over the ``2n`` subsystems ``dims ++ dims`` (rows, then
columns) with only ROW subsystems ever sharded: the column side of every block is complete, so
``U rho U^dagger`` / Kraus need the swap for the row side only.

Reductions never move the state: each rank reduces its block (``pw_sums_strided`` /
``pw_trace_strided`` / ``pw_reduced_from_vec``), ONE all-reduce of ``t`` or ``t^2`` numbers
combines them.  A measurement of a sharded axis collapses on the OWNING ranks (the others
drop their block) and the post state, ``t`` times smaller, is spread over all ranks again.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


@dataclass(frozen=True)
class VAxis:
    g: int        # global subsystem index
    extent: int
    weight: int   # multiplier of this coordinate inside subsystem g's index


class PeerUnavailable(RuntimeError):
    """CUDA IPC peer memory could not be set up on EVERY rank (decided collectively)."""


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

    Construction is COLLECTIVE and all-or-nothing: every rank takes part in both exchanges
    whatever happened locally (a failed allocation travels as a sentinel), the outcome is agreed
    on, and either every rank gets working peer buffers or every rank raises PeerUnavailable
    with its own allocations released -- no rank is left waiting in a collective the others
    skipped, and no rank runs the NCCL swap while its peers run the peer-memory one.
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
        self._opened: List[int] = []
        handles: Optional[List[bytes]] = []
        error = None
        try:
            for _ in range(nbuf):
                ptr = C.c_void_p()
                h = (C.c_ubyte * 64)()
                _lib.check(lib.pw_peer_alloc(numel * self.esize, C.byref(ptr), h), "pw_peer_alloc")
                self.local_ptrs.append(int(ptr.value))
                handles.append(bytes(h))
        except Exception as exc:  # noqa: BLE001 -- reported collectively below
            error, handles = f"rank {rank}: {exc}", None
        gathered: List[Any] = [None] * world
        dist.all_gather_object(gathered, handles)          # unconditional: sentinel None on failure
        if any(g is None for g in gathered):
            self.close()
            raise PeerUnavailable(error or "a peer rank could not allocate / export its shard buffers")
        self.ptrs: List[List[int]] = []  # [buffer][rank]
        try:
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
        except Exception as exc:  # noqa: BLE001
            error = f"rank {rank}: {exc}"
        flags: List[Any] = [None] * world
        dist.all_gather_object(flags, error)                # agree on the outcome of the mapping
        if any(f is not None for f in flags):
            self.close()
            raise PeerUnavailable(next(f for f in flags if f is not None))
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

    # ---- peer memory ------------------------------------------------------------------
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
        """Fused swap + apply; False when the shape is outside the fused kernels."""
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

    def peer_copy(self, dst: torch.Tensor, src_ptr: int, nbytes: int) -> None:
        from . import _lib
        from .core import _arrays as A

        _lib.check(_lib.load().pw_peer_copy(dst.data_ptr(), src_ptr, nbytes, A.stream_ptr()), "pw_peer_copy")

    # ---- apply ------------------------------------------------------------------------
    def apply_vec(self, local: torch.Tensor, op: torch.Tensor, dims, targets, out=None):
        from .core import ffi

        return ffi.apply_vec(local, op, dims, targets, out=out)

    def kraus_axes(self, local: torch.Tensor, ops: torch.Tensor, axes, row_targets, col_targets, out=None):
        from .core import ffi

        return ffi.kraus_axes(local, ops, axes, row_targets, col_targets, out=out)

    def permute(self, local: torch.Tensor, dims, perm, out=None):
        from .core import ffi

        return ffi.permute(local, dims, perm, out=out)

    # ---- reductions -------------------------------------------------------------------
    def norm2(self, local: torch.Tensor) -> torch.Tensor:
        from .core import ffi

        return ffi.norm_stats(local)[:1].clone()

    def sums_strided(self, local, offset, rest, tgt, real_part: bool) -> torch.Tensor:
        from .core import ffi

        return ffi.sums_strided(local, offset, rest, tgt, real_part)

    def probs_from_sums(self, sums: torch.Tensor, cdtype: torch.dtype) -> torch.Tensor:
        from .core import ffi

        return ffi.probs_from_sums(sums, cdtype)

    def draw(self, probs: torch.Tensor, key, cdtype: torch.dtype) -> torch.Tensor:
        from .core import ffi

        return ffi.draw(probs, key, cdtype)

    def trace_strided(self, local, offset, rest, rows, cols) -> torch.Tensor:
        from .core import ffi

        return ffi.trace_strided(local, offset, rest, rows, cols)

    def gather_strided(self, local, offset, axes, probs, outcome, mode: int) -> torch.Tensor:
        from .core import ffi

        return ffi.gather_strided(local, offset, axes, probs, outcome, mode)

    def reduced_from_vec(self, local, dims, keep) -> torch.Tensor:
        from .core import ffi

        return ffi.reduced_from_vec(local, dims, keep)


def rank_digits(rank: int, extents: Sequence[int]) -> List[int]:
    digits = []
    for e in reversed(extents):
        digits.append(rank % e)
        rank //= e
    return list(reversed(digits))


def digits_rank(digits: Sequence[int], extents: Sequence[int]) -> int:
    r = 0
    for d, e in zip(digits, extents):
        r = r * e + d
    return r


def _strides(extents: Sequence[int]) -> List[int]:
    out, acc = [], 1
    for e in reversed(extents):
        out.append(acc)
        acc *= e
    return list(reversed(out))


# ======================================================================================
# common machinery
# ======================================================================================


class ShardedTensor:
    """A tensor over ``global_dims`` block-sharded over ``world`` ranks (virtual-axis layout)."""

    #: subsystems that may be sharded / swapped out (None = all)
    _shardable: Optional[range] = None
    #: whether an operation may be a stack of Kraus operators (K, t, t)
    _stacked_ok = False

    def __init__(self, global_dims: Sequence[int], local: torch.Tensor, vaxes: List[VAxis],
                 sparts: List[VAxis], rank: int, world: int, dist: Any, backend: Any,
                 spare: Optional[torch.Tensor] = None, peers: Optional[PeerBuffers] = None,
                 third: Optional[torch.Tensor] = None):
        self.global_dims = tuple(int(d) for d in global_dims)
        self.local = local.reshape(-1)
        self.vaxes = list(vaxes)
        self.sparts = list(sparts)
        self.rank, self.world = rank, world
        self.dist = dist
        self.backend = backend
        self.spare = spare  # second buffer for out-of-place apply / all-to-all
        self.peers = peers  # IPC-mapped shard buffers: swap = direct NVLink reads, no NCCL
        self.third = third  # receive buffer of the pipelined swap (run_layer)
        self.pipelined_swaps = 0
        self._side = None
        self._flag_side = None
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
    def _layout(cls, global_dims: Sequence[int], plan: Dict[int, int]):
        vaxes, sparts = [], []
        for g, d in enumerate(global_dims):
            f = plan.get(g, 1)
            if f > 1:
                sparts.append(VAxis(g, f, d // f))
            vaxes.append(VAxis(g, d // f, 1))
        return vaxes, sparts

    @classmethod
    def initial_layout(cls, global_dims: Sequence[int], world: int):
        return cls._layout(global_dims, dict(cls.shard_plan(global_dims, world)))

    @staticmethod
    def _own_block(global_dims, full: torch.Tensor, sparts, rank: int) -> torch.Tensor:
        t = full.reshape(tuple(global_dims))
        digits = rank_digits(rank, [s.extent for s in sparts])
        index = [slice(None)] * len(global_dims)
        for s, b in zip(sparts, digits):
            blk = global_dims[s.g] // s.extent
            index[s.g] = slice(b * blk, (b + 1) * blk)
        return t[tuple(index)].clone().reshape(-1)  # never alias the caller's buffer

    # ---- helpers ----------------------------------------------------------------------

    @property
    def local_dims(self) -> Tuple[int, ...]:
        return tuple(v.extent for v in self.vaxes)

    def local_strides(self) -> List[int]:
        return _strides(self.local_dims)

    def my_digits(self) -> List[int]:
        return rank_digits(self.rank, [s.extent for s in self.sparts])

    def sharded_axes(self) -> List[int]:
        return sorted({s.g for s in self.sparts})

    def _parts_of(self, g: int) -> List[int]:
        """Positions of subsystem g's local virtual axes, most significant (largest weight) first."""
        parts = [(v.weight, i) for i, v in enumerate(self.vaxes) if v.g == g and v.extent > 1]
        if not parts:  # a subsystem of dimension 1, or one whose whole index is sharded
            parts = [(v.weight, i) for i, v in enumerate(self.vaxes) if v.g == g][:1]
        parts.sort(reverse=True)
        return [i for _, i in parts]

    def _targets_for(self, axes: Sequence[int]) -> List[int]:
        targets: List[int] = []
        for g in axes:
            pos = self._parts_of(g)
            # the local parts must tile subsystem g's index as a mixed radix number
            expect = 1
            for i in reversed(pos):
                if self.vaxes[i].weight != expect and self.vaxes[i].extent > 1:
                    raise RuntimeError(f"subsystem {g} is not fully local (layout {self.vaxes})")
                expect *= self.vaxes[i].extent
            if expect != self.global_dims[g]:
                raise RuntimeError(f"subsystem {g} is not fully local")
            targets += pos
        return targets

    def _buffers(self) -> torch.Tensor:
        if self.spare is None or self.spare.numel() != self.local.numel():
            if self.peers is not None and self.local.numel() == self.peers.numel:
                # peer mode: every buffer a swap may read must be one of the IPC-exported ones
                self.spare = next(t for t in self.peers.tensors if t.data_ptr() != self.local.data_ptr())
            else:
                self.spare = torch.empty_like(self.local)
        return self.spare

    def _barrier(self) -> None:
        """Stream-ordered cross-rank barrier (a one-element all-reduce; no host sync)."""
        if self._flag is None:
            self._flag = torch.zeros(1, dtype=torch.float32, device=self.local.device)
        self.dist.all_reduce(self._flag)

    def _all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            self.dist.all_reduce(torch.view_as_real(t) if t.is_complex() else t)
        return t

    # ---- argument checks (the seam's error behaviour: core/adapters.py:317-327) -------------

    def _n_subsystems(self) -> int:
        return len(self.global_dims)

    def _check_axes(self, axes: Sequence[int]) -> Tuple[int, ...]:
        axes = tuple(int(a) for a in axes)
        nsub = self._n_subsystems()
        if not axes or len(set(axes)) != len(axes) or any(a < 0 or a >= nsub for a in axes):
            raise ValueError(f"target subsystems {axes} must be distinct indices below {nsub}")
        return axes

    def _validated(self, operator: torch.Tensor, axes: Sequence[int], stacked: bool = False):
        """(operator in the state's dtype, contiguous; axes as a tuple) or the reference's errors:
        ValueError for bad targets, AssertionError for an operator of the wrong shape."""
        axes = self._check_axes(axes)
        t = math.prod(self.global_dims[a] for a in axes)
        shape = tuple(operator.shape)
        if not (shape == (t, t) or (stacked and len(shape) == 3 and shape[1:] == (t, t) and shape[0] >= 1)):
            raise AssertionError(f"Operator shape {shape} does not match expected {(t, t)}")
        if operator.dtype != self.local.dtype:
            operator = operator.to(self.local.dtype)
        return operator.contiguous(), axes

    # ---- axis swap ----------------------------------------------------------------------

    def ensure_local(self, axes: Sequence[int], prefer_shard: Optional[Sequence[int]] = None) -> None:
        if any(s.g in axes for s in self.sparts):
            self.relayout(*self._plan_swap(set(axes), prefer_shard))

    @staticmethod
    def _prime_factors(n: int) -> List[int]:
        out, p = [], 2
        while n > 1:
            while n % p == 0:
                out.append(p)
                n //= p
            p += 1
        return out

    def _choose_parts(self, factors: Sequence[int], needed: set, prefer: Optional[Sequence[int]]) -> List[int]:
        """For every factor (in order) pick a local virtual axis of exactly that extent, splitting
        the top factor off a larger axis when necessary; returns their positions (aligned with
        ``factors``).  Candidates: axes of the preferred subsystems first, then front to back;
        never a needed subsystem, never one outside ``_shardable``."""
        chosen: List[int] = []
        for p in factors:
            order = list(range(len(self.vaxes)))
            if prefer:
                pref = [i for i in order if self.vaxes[i].g in prefer]
                order = pref + [i for i in order if i not in pref]
            pick = None
            for i in order:
                v = self.vaxes[i]
                if i in chosen or v.g in needed or v.extent % p != 0 or v.extent < p:
                    continue
                if self._shardable is not None and v.g not in self._shardable:
                    continue
                pick = i
                break
            if pick is None:
                raise RuntimeError("no admissible axis swap for this operator / world size "
                                   f"(need a local factor {p}; layout {self.vaxes})")
            v = self.vaxes[pick]
            if v.extent != p:  # block split: (e) -> (p, e/p), weights (w*e/p, w)
                hi = VAxis(v.g, p, v.weight * (v.extent // p))
                lo = VAxis(v.g, v.extent // p, v.weight)
                self.vaxes[pick:pick + 1] = [hi, lo]
                chosen = [c + 1 if c > pick else c for c in chosen]
            chosen.append(pick)
        return chosen

    def _plan_swap(self, needed: set, prefer: Optional[Sequence[int]]) -> Tuple[List[int], List[int]]:
        """Which sharded parts become local and which local parts take over their rank-digit
        slots.  First choice: ALL sharded parts (one all-to-all over all ranks; nothing stays
        sharded that a later operator of the layer might need -- the two-layout steady state).
        When the local axes cannot supply the whole world size, only the parts of the needed
        subsystems are localised and the others STAY sharded (an all-to-all inside groups).
        Slots are refined to prime extents first (a pure relabelling), so that a slot can be
        matched factor by factor."""
        saved = (list(self.vaxes), list(self.sparts))
        for everything in (True, False):
            refined: List[VAxis] = []
            slots: List[int] = []
            for s in saved[1]:
                if (everything or s.g in needed) and s.extent > 1:
                    w = s.weight * s.extent
                    for p in self._prime_factors(s.extent):
                        w //= p
                        slots.append(len(refined))
                        refined.append(VAxis(s.g, p, w))
                else:
                    refined.append(s)
            self.vaxes, self.sparts = list(saved[0]), refined
            try:
                positions = self._choose_parts([refined[k].extent for k in slots], needed, prefer)
                return positions, slots
            except RuntimeError:
                if not everything:
                    self.vaxes, self.sparts = saved
                    raise
        raise AssertionError("unreachable")

    def _pack_front(self, positions: List[int]) -> None:
        """Bring the virtual axes at ``positions`` to the front of the buffer (one local pass)."""
        k = len(positions)
        if positions == list(range(k)):
            return
        perm = positions + [i for i in range(len(self.vaxes)) if i not in positions]
        out = self._buffers()
        self.backend.permute(self.local, self.local_dims, perm, out=out)
        self.local, self.spare = out, self.local
        self.vaxes = [self.vaxes[i] for i in perm]
        self.packs += 1

    def relayout(self, positions: Sequence[int], slots: Optional[Sequence[int]] = None, fused=None) -> None:
        """Axis swap: the virtual axes at ``positions`` take over the rank-digit ``slots`` (default:
        all of them), whose sharded parts arrive as the new leading local axes.  Ranks that share
        the digits of the OTHER slots form a group of f = prod(slot extents) members; the swap is
        an all-to-all inside each group (f = world: one all-to-all over all ranks).
        ``fused=(operator, axes)``: apply that operator while gathering (peer mode only)."""
        positions = list(positions)
        slots = list(range(len(self.sparts))) if slots is None else list(slots)
        k = len(positions)
        assert k == len(slots)
        assert [self.vaxes[p].extent for p in positions] == [self.sparts[q].extent for q in slots]
        f = math.prod(self.sparts[q].extent for q in slots)
        self._pack_front(positions)
        recv = self._buffers()
        new_sparts = list(self.sparts)
        for j, q in enumerate(slots):
            new_sparts[q] = self.vaxes[j]
        new_vaxes = [self.sparts[q] for q in slots] + self.vaxes[k:]
        # my group: same digits on the other slots; member c carries digits(c) on `slots`
        ext = [s.extent for s in self.sparts]
        mine = self.my_digits()
        slot_ext = [ext[q] for q in slots]
        members = []
        for c in range(f):
            d = list(mine)
            for q, dq in zip(slots, rank_digits(c, slot_ext)):
                d[q] = dq
            members.append(digits_rank(d, ext))
        c_me = digits_rank([mine[q] for q in slots], slot_ext)
        chunk = self.local.numel() // f
        if self.peers is not None:
            self._barrier()  # every rank has finished writing the buffer we are about to read
            buf = self.peers.index_of(self.local)
            off = c_me * chunk * self.peers.esize
            srcs = [self.peers.ptrs[buf][r] + off for r in members]
            done = False
            if fused is not None:
                op, axes = fused
                self.vaxes, self.sparts = new_vaxes, new_sparts
                done = self.backend.apply_vec_gather(srcs, chunk, recv, op, self.local_dims,
                                                     self._targets_for(axes), rank=c_me)
                if done:
                    self.fused_swaps += 1
            if not done:
                self.backend.gather_chunks(srcs, chunk, recv, rank=c_me)
            self._barrier()  # peers are done reading before anyone overwrites the old buffer
            self.local, self.spare = recv, self.local
            self.vaxes, self.sparts = new_vaxes, new_sparts
            self.swaps += 1
            if fused is not None and not done:
                self._apply_local(fused[0], fused[1])
            return
        # complex buffers travel as (n, 2) reals: one equal piece per group member
        if f == self.world:
            self.dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(self.local))
        else:
            splits = [chunk if r in members else 0 for r in range(self.world)]
            self.dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(self.local),
                                        output_split_sizes=splits, input_split_sizes=splits)
        self.local, self.spare = recv, self.local
        self.vaxes, self.sparts = new_vaxes, new_sparts
        self.swaps += 1
        if fused is not None:
            self._apply_local(fused[0], fused[1])

    # ---- a whole circuit layer with a PIPELINED axis swap ------------------------------------

    def _apply_on(self, x, out, operator, axes, skip: int) -> None:  # pragma: no cover -- overridden
        raise NotImplementedError

    def _third_buffer(self) -> torch.Tensor:
        if self.third is None or self.third.numel() != self.local.numel():
            self.third = torch.empty_like(self.local)
        return self.third

    @staticmethod
    def _stages_keep_program_order(ops, stage: List[int]) -> bool:
        """Reordering by stage is valid iff every pair of operations that share a subsystem keeps
        its program order (operations on disjoint subsystems commute)."""
        for i in range(len(ops)):
            for j in range(i + 1, len(ops)):
                if stage[i] > stage[j] and set(ops[i][1]) & set(ops[j][1]):
                    return False
        return True

    def run_layer(self, ops: Sequence[Tuple[torch.Tensor, Sequence[int]]],
                  next_shard: Optional[Sequence[int]] = None, pipeline: bool = True,
                  marks: Optional[list] = None) -> Dict[str, Any]:
        """Apply ``ops`` = [(operator, subsystems)] (program order) with at most ONE axis swap.

        Operations on disjoint subsystems commute, so the layer is regrouped into
          S0  local operations that touch the OUTGOING subsystems (those about to be sharded),
          S1  the other local operations,
          S2  the operations that need a currently sharded subsystem (after the swap),
        and the swap is hidden behind S1: the outgoing axes lead the buffer, so the piece
        destined to rank c is the contiguous chunk c, itself a dense tensor S1 can be applied to
        on its own.  Rank r works through the chunks in the order r+1, r+2, ..., r; as soon as
        chunk c is final it travels to rank c (a copy-engine pull over NVLink from peer memory
        behind a stream-ordered barrier, or an NCCL send/recv pair) on a side stream while the
        next chunk is being computed.  Every step moves W pieces between W disjoint pairs.
        Falls back to the plain schedule (S0+S1, swap, S2) when the swap needs a pack pass or is
        group-wise, and to operation-by-operation application when the regrouping would break
        the program order.  ``marks``: optional list that receives (label, cuda event) pairs."""
        ops = [self._validated(o, a, stacked=self._stacked_ok) for o, a in ops]
        info: Dict[str, Any] = {"pipelined": False, "s0": 0, "s1": 0, "s2": 0}
        cuda = self.local.is_cuda

        def mark(label):
            if marks is not None and cuda:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((label, ev))

        mark("begin")
        cur = set(self.sharded_axes())
        s2 = [i for i, (_, a) in enumerate(ops) if set(a) & cur]
        if not s2:
            for o, a in ops:
                self._apply_local(o, a)
            info["s1"] = len(ops)
            mark("end")
            return info
        needed = {g for i in s2 for g in ops[i][1]}
        try:
            positions, slots = self._plan_swap(needed, next_shard)
        except RuntimeError:
            # the post-swap operations together touch so many subsystems that ONE swap cannot make
            # them all local: every operation on its own (each needs only its own subsystems local)
            for o, a in ops:
                self.ensure_local(a, next_shard)
                self._apply_local(o, a)
            mark("end")
            return info
        k = len(positions)
        out_g = {self.vaxes[p].g for p in positions}
        stage = [2 if i in s2 else (0 if set(a) & out_g else 1) for i, (_, a) in enumerate(ops)]
        # dependency closure: whatever follows a post-swap operation on a shared subsystem also
        # runs after the swap; whatever precedes an S0 operation on a shared subsystem joins S0
        for i in range(len(ops)):
            for j in range(i + 1, len(ops)):
                if stage[i] == 2 and stage[j] != 2 and set(ops[i][1]) & set(ops[j][1]):
                    stage[j] = 2
        for j in reversed(range(len(ops))):
            for i in range(j):
                if stage[j] == 0 and stage[i] == 1 and set(ops[i][1]) & set(ops[j][1]):
                    stage[i] = 0
        ok = all(not (stage[i] == 2 and set(a) & out_g) for i, (_, a) in enumerate(ops))
        s2 = [i for i in range(len(ops)) if stage[i] == 2]
        if not ok or not self._stages_keep_program_order(ops, stage):
            for o, a in ops:            # no regrouping possible: every operation on its own
                self.ensure_local(a, next_shard)
                self._apply_local(o, a)
            mark("end")
            return info
        st0 = [ops[i] for i in range(len(ops)) if stage[i] == 0]
        st1 = [ops[i] for i in range(len(ops)) if stage[i] == 1]
        st2 = [ops[i] for i in range(len(ops)) if stage[i] == 2]
        info.update(s0=len(st0), s1=len(st1), s2=len(st2))
        W = self.world
        full = len(slots) == len(self.sparts) and math.prod(self.sparts[q].extent for q in slots) == W
        can_pipe = (pipeline and W > 1 and full and positions == list(range(k)) and len(st1) > 0
                    and (self.peers is None or len(self.peers.local_ptrs) >= 3))
        if not can_pipe:
            for o, a in st0 + st1:
                self._apply_local(o, a)
            mark("local_done")
            self.relayout(positions, slots)
            mark("swap_done")
            for o, a in st2:
                self._apply_local(o, a)
            mark("end")
            return info

        for o, a in st0:
            self._apply_local(o, a)
        mark("s0_done")
        X, Y = self.local, self._buffers()
        if self.peers is not None:
            others = [t for t in self.peers.tensors if t.data_ptr() not in (X.data_ptr(), Y.data_ptr())]
            Z = others[0]
        else:
            Z = self._third_buffer()
        chunk = X.numel() // W
        src_buf = X if len(st1) % 2 == 0 else Y
        side = None
        if cuda:
            if self._side is None:
                self._side = torch.cuda.Stream(priority=-1)
            side = self._side
            side.wait_stream(torch.cuda.current_stream())
        r = self.rank
        ranks_by_digit = list(range(W))  # full swap: group member c is rank c
        for step in range(W):
            c = (r + 1 + step) % W
            a, b = X[c * chunk:(c + 1) * chunk], Y[c * chunk:(c + 1) * chunk]
            p = (r - step - 1) % W
            dst = Z[p * chunk:(p + 1) * chunk]
            for j, (o, ax) in enumerate(st1):
                # the own chunk (last step) never travels: its last operation writes it in place
                out = dst if (c == r and j == len(st1) - 1) else b
                self._apply_on(a, out, o, ax, k)
                a, b = b, a
            if c != r:
                self._move_piece(a, ranks_by_digit[c], dst, ranks_by_digit[p], src_buf, chunk, side)
        mark("s1_done")
        if cuda:
            torch.cuda.current_stream().wait_stream(side)
        if self.peers is not None:
            self._barrier()  # every rank has pulled its pieces before anyone overwrites the source
        mark("swap_done")
        new_sparts = list(self.sparts)
        for j, q in enumerate(slots):
            new_sparts[q] = self.vaxes[j]
        self.vaxes = [self.sparts[q] for q in slots] + self.vaxes[k:]
        self.sparts = new_sparts
        keep = X if src_buf is Y else Y
        self.local, self.spare = Z, keep
        if self.peers is None:
            self.third = src_buf
        self.swaps += 1
        self.pipelined_swaps += 1
        info["pipelined"] = True
        for o, a in st2:
            self._apply_local(o, a)
        mark("end")
        return info

    def _move_piece(self, piece, dst_rank: int, dst, src_rank: int, src_buf, chunk: int, side) -> None:
        """One step of the pipelined swap: `piece` (a finished chunk of this rank) goes to
        ``dst_rank``; ``dst`` receives the chunk this rank needs from ``src_rank``."""
        cuda = piece.is_cuda
        if self.peers is not None:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(side):
                side.wait_event(ev)
                if self._flag_side is None:
                    self._flag_side = torch.zeros(1, dtype=torch.float32, device=piece.device)
                self.dist.all_reduce(self._flag_side)   # every rank has finished this step's chunk
                buf = self.peers.index_of(src_buf)
                src = self.peers.ptrs[buf][src_rank] + self.rank * chunk * self.peers.esize
                self.backend.peer_copy(dst, src, chunk * self.peers.esize)
            return
        reqs = [self.dist.P2POp(self.dist.isend, torch.view_as_real(piece), dst_rank),
                self.dist.P2POp(self.dist.irecv, torch.view_as_real(dst), src_rank)]
        if cuda:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(side):
                side.wait_event(ev)
                for req in self.dist.batch_isend_irecv(reqs):
                    req.wait()
        else:
            for req in self.dist.batch_isend_irecv(reqs):
                req.wait()

    def _apply_local(self, operator, axes) -> None:  # pragma: no cover -- overridden
        raise NotImplementedError

    # ---- measurement plumbing shared by psi and rho ----------------------------------------

    def _global_positions(self, axes: Sequence[int], parts: List[int]) -> np.ndarray:
        """For every LOCAL joint index over the virtual axes ``parts`` (row-major, in that order):
        the position in the global joint outcome (mixed radix over ``axes`` in the given order,
        core/kernels.py:28-41), with this rank's digits filled in for the sharded parts."""
        mult, acc = {}, 1
        for g in reversed(list(axes)):
            mult[g] = acc
            acc *= self.global_dims[g]
        pos = np.zeros((1,), dtype=np.int64)
        for i in parts:
            v = self.vaxes[i]
            pos = (pos[:, None] + (np.arange(v.extent, dtype=np.int64) * v.weight * mult[v.g])[None, :]).reshape(-1)
        const = 0
        for s, b in zip(self.sparts, self.my_digits()):
            if s.g in mult:
                const += b * s.weight * mult[s.g]
        return pos + const

    def _combine_probs(self, local_sums: torch.Tensor, axes, parts: List[int]) -> torch.Tensor:
        """Scatter this rank's partial sums into the global outcome vector, ONE all-reduce,
        normalise (core/kernels.py:40, :58)."""
        t = math.prod(self.global_dims[g] for g in axes)
        full = torch.zeros(t, dtype=torch.float64, device=local_sums.device)
        pos = torch.as_tensor(self._global_positions(axes, parts), device=local_sums.device)
        full.index_copy_(0, pos, local_sums.reshape(-1))
        self._all_reduce(full)
        return self.backend.probs_from_sums(full, self.local.dtype)

    def _outcome_digits(self, outcome: int, axes: Sequence[int]) -> Dict[int, int]:
        idx, rem = {}, int(outcome)
        for g in reversed(list(axes)):
            idx[g] = rem % self.global_dims[g]
            rem //= self.global_dims[g]
        return idx

    def _owner_and_offset(self, idx: Dict[int, int], strides: Dict[int, int]) -> Tuple[bool, int]:
        """Does this rank hold outcome ``idx`` (its digits on the sharded measured parts match),
        and the element offset of that outcome inside the local block."""
        own = True
        for s, b in zip(self.sparts, self.my_digits()):
            if s.g in idx and (idx[s.g] // s.weight) % s.extent != b:
                own = False
        off = 0
        for i, v in enumerate(self.vaxes):
            if v.g in idx and i in strides:
                off += ((idx[v.g] // v.weight) % v.extent) * strides[i]
        return own, off

    def _check_respreadable(self, post_vaxes: List[VAxis], measured: Sequence[int]) -> None:
        """Raise BEFORE any collective when the post-measurement state cannot be spread over all
        ranks again (its remaining local axes cannot supply the factor the measured sharded parts
        held) -- every rank sees the same layout, so every rank raises the same error."""
        f = math.prod(s.extent for s in self.sparts if s.g in measured)
        if f == 1:
            return
        probe = ShardedTensor.__new__(type(self))
        probe.vaxes, probe._shardable = list(post_vaxes), self._shardable
        try:
            ShardedTensor._choose_parts(probe, self._prime_factors(f), set(), None)
        except RuntimeError:
            raise ValueError(
                f"measuring subsystems {tuple(measured)} leaves a state whose local axes "
                f"{[(v.g, v.extent) for v in post_vaxes]} cannot be spread over the {f} ranks that shared "
                "the measured axes; measure on fewer ranks or gather the state first") from None

    def _respread(self, post: Optional[torch.Tensor], post_vaxes: List[VAxis], measured: Sequence[int],
                  idx: Dict[int, int]) -> Tuple[torch.Tensor, List[VAxis], List[VAxis]]:
        """After a collapse on the owning ranks: spread the post state over ALL ranks again.

        The measured sharded parts (factor f of the world size) no longer exist; the owners cut
        their block along f fresh leading parts and send one piece to each rank of their group.
        The post state is t times smaller than the state, so this moves little data.
        Returns (local, vaxes, sparts) of the new layout."""
        dtype, device = self.local.dtype, self.local.device
        old_ext = [s.extent for s in self.sparts]
        mine = self.my_digits()
        kept = [k for k, s in enumerate(self.sparts) if s.g not in measured]
        f = math.prod(s.extent for s in self.sparts if s.g in measured)
        if f == 1:
            return post, post_vaxes, [self.sparts[k] for k in kept]
        # fresh parts: computed identically on every rank from the layout alone
        probe = ShardedTensor.__new__(type(self))
        probe.vaxes, probe._shardable = list(post_vaxes), self._shardable
        positions = ShardedTensor._choose_parts(probe, self._prime_factors(f), set(), None)
        new_vaxes = probe.vaxes
        perm = positions + [i for i in range(len(new_vaxes)) if i not in positions]
        fresh = [new_vaxes[i] for i in positions]
        local_vaxes = [new_vaxes[i] for i in perm[len(positions):]]
        n_piece = math.prod(v.extent for v in local_vaxes)
        recv = torch.empty(n_piece, dtype=dtype, device=device)
        new_sparts = [self.sparts[k] for k in kept] + fresh
        new_ext = [s.extent for s in new_sparts]
        fresh_ext = [v.extent for v in fresh]
        ops = []
        pieces = None
        if post is not None:  # owner: cut the block along the fresh parts (front of the buffer)
            dims = [v.extent for v in new_vaxes]
            pieces = self.backend.permute(post.reshape(-1), dims, perm).reshape(f, n_piece)
            for c in range(f):
                dst = digits_rank([mine[k] for k in kept] + rank_digits(c, fresh_ext), new_ext)
                if dst == self.rank:
                    recv.copy_(pieces[c])
                else:
                    ops.append(self.dist.P2POp(self.dist.isend, torch.view_as_real(pieces[c]), dst))
        # every rank receives exactly one piece, from the owner of its group
        my_new = rank_digits(self.rank, new_ext)
        src_digits = list(mine)
        for j, k in enumerate(kept):
            src_digits[k] = my_new[j]
        for k, s in enumerate(self.sparts):
            if s.g in measured:
                src_digits[k] = (idx[s.g] // s.weight) % s.extent  # the owner's digit on that part
        src = digits_rank(src_digits, old_ext)
        if src != self.rank:
            ops.append(self.dist.P2POp(self.dist.irecv, torch.view_as_real(recv), src))
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()
        return recv, local_vaxes, new_sparts

    # ---- host buffers (one pinned shard per rank) --------------------------------------------

    def load_host(self, host: torch.Tensor) -> None:
        """Replace this rank's shard (current layout) with a pinned host shard, asynchronously on
        the current stream: every rank uploads over its own PCIe link."""
        if host.numel() != self.local.numel() or host.dtype != self.local.dtype:
            raise ValueError("host shard must match the local shard's size and dtype")
        self.local.reshape(-1).copy_(host.reshape(-1), non_blocking=True)

    def store_host(self, host: torch.Tensor) -> None:
        """Copy this rank's shard (current layout: `vaxes`, `sparts`) into a pinned host shard."""
        if host.numel() != self.local.numel() or host.dtype != self.local.dtype:
            raise ValueError("host shard must match the local shard's size and dtype")
        host.reshape(-1).copy_(self.local.reshape(-1), non_blocking=True)

    # ---- reassembly (tests / preflight) -----------------------------------------------------

    def gather_full(self) -> torch.Tensor:
        """Reassemble the full tensor in canonical subsystem order on every rank (tests)."""
        pieces = [torch.empty_like(self.local) for _ in range(self.world)]
        if self.world > 1:
            self.dist.all_gather([torch.view_as_real(p) for p in pieces],
                                 torch.view_as_real(self.local.contiguous()))
        else:
            pieces = [self.local]
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


# ======================================================================================
# state vector
# ======================================================================================


class ShardedStateVector(ShardedTensor):
    """A state vector of ``global_dims`` block-sharded over ``world`` ranks."""

    @classmethod
    def from_full(cls, global_dims, full: torch.Tensor, rank: int, world: int, dist, backend):
        """Every rank holds the full vector (tests / small cases): keep only the own block."""
        vaxes, sparts = cls.initial_layout(global_dims, world)
        local = cls._own_block(global_dims, full, sparts, rank)
        return cls(global_dims, local, vaxes, sparts, rank, world, dist, backend)

    # ---- apply (core/kernels.py:62-98) ------------------------------------------------------

    def apply_operation(self, operator: torch.Tensor, axes: Sequence[int],
                        prefer_shard: Optional[Sequence[int]] = None) -> None:
        """psi <- (operator on global subsystems ``axes``) psi; swaps axes in if needed."""
        operator, axes = self._validated(operator, axes)
        if self.peers is not None and any(s.g in axes for s in self.sparts):
            # fused: the operator is applied while the swapped data streams in over NVLink
            positions, slots = self._plan_swap(set(axes), prefer_shard)
            self.relayout(positions, slots, fused=(operator, axes))
            return
        self.ensure_local(axes, prefer_shard)
        self._apply_local(operator, axes)

    def _apply_local(self, operator, axes) -> None:
        out = self._buffers()
        self._apply_on(self.local, out, operator, axes, 0)
        self.local, self.spare = out, self.local

    def _apply_on(self, x, out, operator, axes, skip: int) -> None:
        targets = [t - skip for t in self._targets_for(axes)]
        assert min(targets) >= 0
        self.backend.apply_vec(x, operator, self.local_dims[skip:], targets, out=out)

    # ---- reductions ---------------------------------------------------------------------------

    def norm2(self) -> torch.Tensor:
        """sum |psi|^2 over all ranks (device tensor)."""
        return self._all_reduce(self.backend.norm2(self.local))

    def _split(self, axes: Sequence[int]):
        parts = [i for g in axes for i in self._parts_of(g)]
        rest = [i for i, v in enumerate(self.vaxes) if v.g not in axes]
        return parts, rest

    def probs(self, axes: Sequence[int]) -> torch.Tensor:
        """Joint outcome probabilities of subsystems ``axes`` (given order), identical on every
        rank; no data movement whether or not the axes are sharded (core/kernels.py:28-41)."""
        axes = self._check_axes(axes)
        parts, rest = self._split(axes)
        st, ext = self.local_strides(), self.local_dims
        sums = self.backend.sums_strided(self.local, 0, [(ext[i], st[i]) for i in rest],
                                         [(ext[i], st[i]) for i in parts], False)
        return self._combine_probs(sums, axes, parts)

    def reduced(self, keep: Sequence[int], prefer_shard: Optional[Sequence[int]] = None) -> torch.Tensor:
        """Tr_rest |psi><psi| for the subsystems ``keep`` (given order), on every rank
        (state/utils/trace_out.py:17-25): local partial + one all-reduce of t^2 numbers."""
        keep = self._check_axes(keep)
        self.ensure_local(keep, prefer_shard)
        part = self.backend.reduced_from_vec(self.local, self.local_dims, self._targets_for(keep))
        return self._all_reduce(part)

    def measure(self, axes: Sequence[int], key) -> Tuple[int, "ShardedStateVector", torch.Tensor]:
        """Projective measurement of ``axes`` (core/kernels.py:189-206): returns (outcome, post
        state over the remaining subsystems, probabilities).  Every rank draws from the same
        all-reduced probabilities with the same key, so the outcome is identical everywhere;
        the collapse happens on the ranks that own the outcome's sharded digits."""
        axes = self._check_axes(axes)
        axes = list(axes)
        parts, rest = self._split(axes)
        self._check_respreadable([self.vaxes[i] for i in rest], axes)
        probs = self.probs(axes)
        outcome_t = self.backend.draw(probs, key, self.local.dtype)
        outcome = int(outcome_t.item())
        idx = self._outcome_digits(outcome, axes)
        st, ext = self.local_strides(), self.local_dims
        own, off = self._owner_and_offset(idx, {i: st[i] for i in parts})
        post = None
        if own:
            post = self.backend.gather_strided(self.local, off, [(ext[i], st[i]) for i in rest],
                                               probs, outcome_t, 1)
        post_vaxes = [self.vaxes[i] for i in rest]
        local, vaxes, sparts = self._respread(post, post_vaxes, axes, idx)
        # renumber the remaining subsystems
        remaining = [g for g in range(len(self.global_dims)) if g not in axes]
        new_id = {g: k for k, g in enumerate(remaining)}
        vaxes = [VAxis(new_id[v.g], v.extent, v.weight) for v in vaxes]
        sparts = [VAxis(new_id[s.g], s.extent, s.weight) for s in sparts]
        dims = [self.global_dims[g] for g in remaining]
        out = ShardedStateVector(dims, local, vaxes, sparts, self.rank, self.world, self.dist, self.backend)
        return outcome, out, probs


# ======================================================================================
# density matrix
# ======================================================================================


class ShardedDensityMatrix(ShardedTensor):
    """A density matrix over subsystems ``dims``, its ROW side block-sharded over ``world`` ranks.

    Internally a tensor over the 2n subsystems ``dims ++ dims`` (row g = subsystem g, column g =
    subsystem n + g) in which only row subsystems are ever sharded or swapped, so every block is
    a set of complete rows: ``local`` is a row-major (local rows x N) matrix."""

    def __init__(self, dims: Sequence[int], local, vaxes, sparts, rank, world, dist, backend,
                 spare=None, peers=None, third=None):
        self.dims = tuple(int(d) for d in dims)
        self.n = len(self.dims)
        self._shardable = range(self.n)
        super().__init__(self.dims + self.dims, local, vaxes, sparts, rank, world, dist, backend,
                         spare=spare, peers=peers, third=third)

    _stacked_ok = True

    def _n_subsystems(self) -> int:
        return self.n

    @classmethod
    def initial_layout(cls, dims: Sequence[int], world: int):
        dims = tuple(dims)
        plan = dict(cls.shard_plan(dims, world))  # row subsystems only
        return cls._layout(dims + dims, plan)

    @classmethod
    def from_full(cls, dims, full: torch.Tensor, rank: int, world: int, dist, backend):
        dims = tuple(dims)
        vaxes, sparts = cls.initial_layout(dims, world)
        local = cls._own_block(dims + dims, full, sparts, rank)
        return cls(dims, local, vaxes, sparts, rank, world, dist, backend)

    # ---- geometry -----------------------------------------------------------------------------

    def _col_stride(self) -> Dict[int, int]:
        """Element stride of column subsystem g (columns are whole, unsplit axes)."""
        st = self.local_strides()
        out = {}
        for i, v in enumerate(self.vaxes):
            if v.g >= self.n:
                assert v.weight == 1 and v.extent == self.dims[v.g - self.n]
                out[v.g - self.n] = st[i]
        return out

    def _diag(self):
        """(diag stride of every row virtual axis, constant offset): walking a row part by one
        step also moves the matching column subsystem by ``weight``, which stays on the diagonal
        of rho; the rank's digits fix the column offset of the sharded parts."""
        st, cs = self.local_strides(), self._col_stride()
        dstride = {i: st[i] + v.weight * cs[v.g] for i, v in enumerate(self.vaxes) if v.g < self.n}
        const = sum(b * s.weight * cs[s.g] for s, b in zip(self.sparts, self.my_digits()))
        return dstride, const

    # ---- apply / Kraus (core/kernels.py:101-186) ---------------------------------------------------

    def apply_kraus(self, operators: torch.Tensor, axes: Sequence[int],
                    prefer_shard: Optional[Sequence[int]] = None) -> None:
        """rho <- sum_k K_k rho K_k^dagger on subsystems ``axes``; ONE HBM pass per block.  Only
        the row side can be sharded, so at most one axis swap (of rows) precedes it."""
        operators, axes = self._validated(operators, axes, stacked=True)
        self.ensure_local(axes, prefer_shard)
        self._apply_local(operators, axes)

    def apply_operation(self, operator: torch.Tensor, axes: Sequence[int],
                        prefer_shard: Optional[Sequence[int]] = None) -> None:
        """rho <- U rho U^dagger."""
        self.apply_kraus(operator.reshape((1,) + tuple(operator.shape[-2:])), axes, prefer_shard)

    def _apply_local(self, operators, axes) -> None:
        out = self._buffers()
        self._apply_on(self.local, out, operators, axes, 0)
        self.local, self.spare = out, self.local

    def _apply_on(self, x, out, operators, axes, skip: int) -> None:
        if operators.dim() == 2:
            operators = operators.reshape((1,) + tuple(operators.shape))
        rows = [t - skip for t in self._targets_for(axes)]
        cols = [t - skip for t in self._targets_for([self.n + a for a in axes])]
        assert min(rows) >= 0
        self.backend.kraus_axes(x, operators, self.local_dims[skip:], rows, cols, out=out)

    # ---- reductions -----------------------------------------------------------------------------

    def _row_split(self, axes: Sequence[int]):
        parts = [i for g in axes for i in self._parts_of(g)]
        rest = [i for i, v in enumerate(self.vaxes) if v.g < self.n and v.g not in axes]
        return parts, rest

    def trace(self) -> torch.Tensor:
        """Tr rho (device float64 tensor[1]) over all ranks."""
        ds, const = self._diag()
        ext = self.local_dims
        rows = [i for i, v in enumerate(self.vaxes) if v.g < self.n]
        sums = self.backend.sums_strided(self.local, const, [(ext[i], ds[i]) for i in rows], [], True)
        return self._all_reduce(sums)

    def probs(self, axes: Sequence[int]) -> torch.Tensor:
        """Joint outcome probabilities (core/kernels.py:44-59): each rank sums the diagonal
        entries of its own rows, ONE all-reduce of t numbers."""
        axes = self._check_axes(axes)
        ds, const = self._diag()
        ext = self.local_dims
        parts, rest = self._row_split(axes)
        sums = self.backend.sums_strided(self.local, const, [(ext[i], ds[i]) for i in rest],
                                         [(ext[i], ds[i]) for i in parts], True)
        return self._combine_probs(sums, axes, parts)

    def trace_out(self, keep: Sequence[int], prefer_shard: Optional[Sequence[int]] = None) -> torch.Tensor:
        """Tr_rest rho for the subsystems ``keep`` (given order) on every rank
        (core/kernels.py:301-319): local partial over the block's rows + one all-reduce of t^2."""
        keep = self._check_axes(keep)
        self.ensure_local(keep, prefer_shard)
        ds, const = self._diag()
        st, ext, cs = self.local_strides(), self.local_dims, self._col_stride()
        parts, rest = self._row_split(keep)
        part = self.backend.trace_strided(
            self.local, const, [(ext[i], ds[i]) for i in rest], [(ext[i], st[i]) for i in parts],
            [(self.dims[g], cs[g]) for g in keep])
        return self._all_reduce(part)

    def measure(self, axes: Sequence[int], key) -> Tuple[int, "ShardedDensityMatrix", torch.Tensor]:
        """Projective measurement (core/kernels.py:209-227): post = rho[o,:,o,:] / (p_o + 1e-18)
        taken on the ranks that own outcome o's rows, then spread over all ranks again."""
        axes = self._check_axes(axes)
        axes = list(axes)
        parts, rest = self._row_split(axes)
        rest_cols = [g for g in range(self.n) if g not in axes]
        self._check_respreadable([self.vaxes[i] for i in rest] +
                                 [VAxis(self.n + g, self.dims[g], 1) for g in rest_cols], axes)
        probs = self.probs(axes)
        outcome_t = self.backend.draw(probs, key, self.local.dtype)
        outcome = int(outcome_t.item())
        idx = self._outcome_digits(outcome, axes)
        st, ext, cs = self.local_strides(), self.local_dims, self._col_stride()
        own, off = self._owner_and_offset(idx, {i: st[i] for i in parts})
        off += sum(idx[g] * cs[g] for g in axes)
        post = None
        if own:
            view = [(ext[i], st[i]) for i in rest] + [(self.dims[g], cs[g]) for g in rest_cols]
            post = self.backend.gather_strided(self.local, off, view, probs, outcome_t, 2)
        new_id = {g: k for k, g in enumerate(rest_cols)}
        m = len(rest_cols)
        post_vaxes = [self.vaxes[i] for i in rest] + [VAxis(self.n + g, self.dims[g], 1) for g in rest_cols]
        local, vaxes, sparts = self._respread(post, post_vaxes, axes, idx)

        def ren(v: VAxis) -> VAxis:
            return VAxis(new_id[v.g] if v.g < self.n else m + new_id[v.g - self.n], v.extent, v.weight)

        dims = [self.dims[g] for g in rest_cols]
        out = ShardedDensityMatrix(dims, local, [ren(v) for v in vaxes], [ren(s) for s in sparts],
                                   self.rank, self.world, self.dist, self.backend)
        return outcome, out, probs

    def gather_full(self) -> torch.Tensor:
        n_tot = math.prod(self.dims)
        return super().gather_full().reshape(n_tot, n_tot)
