"""
Multi-rank logic of the sharded state vector (axis bookkeeping, all-to-all axis swap,
pack, reductions) on CPU with the gloo backend, world sizes 2 and 4.  The local
arithmetic is injected (NumPy oracle here, libpwb200 on GPUs), so this exercises
exactly the host-side code the N>1 GPU path runs.
"""

import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import adapters_np as oad
from oracle import operators as ops
from photon_weave_b200.sharded import ShardedStateVector


class NumpyBackend:
    def apply_vec(self, local, op, dims, targets, out=None):
        res = oad.apply_operation_vector(tuple(dims), tuple(targets), local.numpy().reshape(-1, 1),
                                         op.numpy(), True)
        res = torch.from_numpy(np.ascontiguousarray(res.reshape(-1)))
        if out is not None:
            out.copy_(res)
            return out
        return res

    def permute(self, local, dims, perm, out=None):
        res = torch.from_numpy(
            np.ascontiguousarray(np.transpose(local.numpy().reshape(dims), perm)).reshape(-1))
        if out is not None:
            out.copy_(res)
            return out
        return res

    def norm2(self, local):
        return torch.tensor([float(np.sum(np.abs(local.numpy()) ** 2))], dtype=torch.float64)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, dims, program, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N = math.prod(dims)
        full = torch.from_numpy(ops.random_state(N, 5).reshape(-1))
        st = ShardedStateVector.from_full(dims, full, rank, world, dist, NumpyBackend())
        ref = full.numpy().reshape(-1, 1)
        for axes, seed, prefer in program:
            t = math.prod(dims[a] for a in axes)
            op = ops.random_unitary(t, seed)
            st.apply_operation(torch.from_numpy(op), axes, prefer_shard=prefer)
            ref = oad.apply_operation_vector(dims, axes, ref, op, True)
        got = st.gather_full().numpy()
        err = float(np.abs(got - ref.reshape(-1)).max())
        nrm = float(st.norm2().item())
        if rank == 0:
            results.put((err, nrm, st.swaps, st.packs))
    finally:
        dist.destroy_process_group()


def _run(world, dims, program):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dims, program, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    return q.get()


def test_shard_plan_factors_world_over_leading_axes():
    assert ShardedStateVector.shard_plan((6,) * 12, 8) == [(0, 2), (1, 2), (2, 2)]
    assert ShardedStateVector.shard_plan((6,) * 12, 2) == [(0, 2)]
    assert ShardedStateVector.shard_plan((8, 8, 8), 8) == [(0, 8)]
    with pytest.raises(ValueError):
        ShardedStateVector.shard_plan((5, 5), 2)


def test_local_ops_need_no_swap_world2():
    dims = (4, 3, 2, 3)
    err, nrm, swaps, packs = _run(2, dims, [((1,), 1, None), ((3, 2), 2, None), ((2,), 3, None)])
    assert err < 1e-13 and abs(nrm - 1) < 1e-12 and swaps == 0 and packs == 0


def test_axis_swap_when_operator_touches_sharded_axis_world2():
    dims = (4, 3, 2, 6)
    program = [((0,), 1, None), ((1,), 2, None), ((0, 3), 3, None), ((3,), 4, None), ((2, 0), 5, None)]
    err, nrm, swaps, packs = _run(2, dims, program)
    assert err < 1e-13 and abs(nrm - 1) < 1e-12 and swaps >= 1


def test_two_alternating_layouts_need_no_pack_in_steady_state_world4():
    dims = (6, 6, 2, 6, 6)
    # layout A shards axes 0,1 (factor 2 each); B shards axes 3,4
    program = []
    for rep in range(3):
        program += [((2,), 10 + rep, None), ((3, 4), 20 + rep, None)]         # local under A
        program += [((0,), 30 + rep, (3, 4)), ((1, 2), 40 + rep, (3, 4))]      # needs A -> B
        program += [((3,), 50 + rep, (0, 1)), ((4,), 60 + rep, (0, 1))]        # needs B -> A
    err, nrm, swaps, packs = _run(4, dims, program)
    assert err < 1e-12 and abs(nrm - 1) < 1e-12
    assert swaps == 6
    assert packs <= 2  # only the first A->B (splitting axes 3,4) may pack


def test_bench_layer_schedule_world2():
    """The layer schedule bench.py --gpus N uses: every op once, one swap per layer."""
    dims = (6, 6, 6, 6)
    program = [((m,), m + 1, None) for m in range(4)]
    program += [((0, 1), 7, (2,)), ((2, 3), 8, (0,))]
    err, nrm, swaps, packs = _run(2, dims, program)
    assert err < 1e-13 and abs(nrm - 1) < 1e-12
