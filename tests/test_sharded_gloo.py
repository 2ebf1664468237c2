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
from oracle import rng_threefry as orng
from photon_weave_b200.sharded import ShardedDensityMatrix, ShardedStateVector


def _strided(local, offset, pairs):
    a = local.numpy()
    ext = tuple(int(e) for e, _ in pairs)
    st = tuple(int(s) * a.itemsize for _, s in pairs)
    return np.lib.stride_tricks.as_strided(a[offset:], shape=ext, strides=st, writeable=False)


class NumpyBackend:
    """The libpwb200 entry points the multi-GPU layer calls, restated with the NumPy oracle."""

    def kraus_axes(self, local, ops, axes, row_targets, col_targets, out=None):
        x = local.numpy().reshape(-1, 1)
        acc = 0
        for k in ops.numpy():
            y = oad.apply_operation_vector(tuple(axes), tuple(row_targets), x, k, True)
            acc = acc + oad.apply_operation_vector(tuple(axes), tuple(col_targets), y, k.conj(), True)
        res = torch.from_numpy(np.ascontiguousarray(acc.reshape(-1)))
        if out is not None:
            out.copy_(res)
            return out
        return res

    def sums_strided(self, local, offset, rest, tgt, real_part):
        v = _strided(local, offset, list(tgt) + list(rest))
        w = v.real if real_part else np.abs(v) ** 2
        t = math.prod(e for e, _ in tgt) if tgt else 1
        return torch.from_numpy(np.ascontiguousarray(w.reshape(t, -1).sum(axis=1)).astype(np.float64))

    def probs_from_sums(self, sums, cdtype):
        s = sums.numpy()
        return torch.from_numpy(s / s.sum())

    def draw(self, probs, key, cdtype):
        p = probs.numpy()
        return torch.tensor(min(orng.choice(key, p.shape[0], p), p.shape[0] - 1), dtype=torch.int64)

    def trace_strided(self, local, offset, rest, rows, cols):
        v = _strided(local, offset, list(rows) + list(cols) + list(rest))
        tr = math.prod(e for e, _ in rows) if rows else 1
        tc = math.prod(e for e, _ in cols) if cols else 1
        return torch.from_numpy(np.ascontiguousarray(v.reshape(tr, tc, -1).sum(axis=2)))

    def gather_strided(self, local, offset, view, probs, outcome, mode):
        v = np.ascontiguousarray(_strided(local, offset, view)).reshape(-1)
        if mode:
            p = probs.numpy()[int(outcome)] + 1e-18
            v = v / (np.sqrt(p) if mode == 1 else p)
        return torch.from_numpy(v)

    def reduced_from_vec(self, local, dims, keep):
        n = len(dims)
        rest = [i for i in range(n) if i not in keep]
        a = local.numpy().reshape(dims).transpose(list(keep) + rest)
        a = a.reshape(math.prod(dims[i] for i in keep), -1)
        return torch.from_numpy(np.ascontiguousarray(a @ a.conj().T))

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
        # host round trip of the shard in its current layout (the e2e leg of bench.py --gpus N)
        host = torch.empty(st.local.numel(), dtype=st.local.dtype)
        st.store_host(host)
        st.local.zero_()
        st.load_host(host)
        with pytest.raises(ValueError):
            st.load_host(host[:-1])
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


# ---- reductions / measurement of a sharded state vector ---------------------------------------


def _vec_reductions_worker(rank, world, port, dims, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N = math.prod(dims)
        psi = ops.random_state(N, 11)
        st = ShardedStateVector.from_full(dims, torch.from_numpy(psi.reshape(-1)), rank, world, dist,
                                          NumpyBackend())
        # one swap first so that split virtual axes and non-leading shards are exercised too
        u = ops.random_unitary(dims[0], 3)
        st.apply_operation(torch.from_numpy(u), (0,))
        ref = oad.apply_operation_vector(dims, (0,), psi, u, True)
        errs = []
        n = len(dims)
        for axes in [(0,), (1,), (n - 1, 0), (1, 2, 0), tuple(range(n))]:
            p = st.probs(axes).numpy()
            _, _, p_ref, _ = oad.measure_vector_with_probs(dims, axes, ref, orng.PRNGKey(0))
            errs.append(np.abs(p - p_ref).max())
        for keep in [(1,), (0, n - 1), (2, 0)]:
            r = st.reduced(keep).numpy()
            errs.append(np.abs(r - oad.trace_out_vector(dims, keep, ref, False)).max())
        for seed, axes in enumerate([(0,), (n - 1, 1), (1, 0), (2,)]):
            key = orng.PRNGKey(20 + seed)
            o, post, probs = st.measure(axes, key)
            o_ref, post_ref, _ = oad.measure_vector(dims, axes, ref, key)
            assert o == int(o_ref), (axes, o, int(o_ref))
            got = post.gather_full().numpy()
            errs.append(np.abs(got - post_ref.reshape(-1)).max())
            assert abs(float(post.norm2().item()) - 1.0) < 1e-9
            # the post state is spread over ALL ranks again
            assert post.local.numel() * world == got.size
        if rank == 0:
            results.put(max(errs))
    finally:
        dist.destroy_process_group()


def _spawn(world, worker, *args):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port) + args + (q,)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    return q.get()


@pytest.mark.parametrize("world,dims", [(2, (4, 3, 2, 6)), (4, (6, 6, 2, 4, 2)), (4, (6, 6, 2, 4))])
def test_sharded_vector_probs_reduced_measure(world, dims):
    assert _spawn(world, _vec_reductions_worker, dims) < 1e-12


# ---- sharded density matrix: apply, Kraus, probs, partial trace, measurement -----------------


def _rho_worker(rank, world, port, dims, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = len(dims)
        N = math.prod(dims)
        rho = ops.random_density(N, 5)
        sd = ShardedDensityMatrix.from_full(dims, torch.from_numpy(rho.reshape(-1)), rank, world, dist,
                                            NumpyBackend())
        assert sd.local.numel() * world == N * N
        ref = rho
        errs = []
        program = [("u", (1,), 1), ("k", (0,), 2), ("u", (0, n - 1), 3), ("k", (2,), 4), ("u", (1, 0), 5),
                   ("k", (n - 1,), 6)]
        for kind, axes, seed in program:
            t = math.prod(dims[a] for a in axes)
            if kind == "u":
                op = ops.random_unitary(t, seed)
                sd.apply_operation(torch.from_numpy(op), axes)
                ref = oad.apply_operation_matrix(dims, axes, ref, op, True)
            else:
                ks = ops.loss_channel(t, 0.15)
                sd.apply_kraus(torch.from_numpy(ks), axes)
                ref = oad.apply_kraus_matrix(dims, axes, ref, ks, True)
            errs.append(np.abs(sd.gather_full().numpy() - ref).max())
            # only row subsystems are ever sharded
            assert all(s.g < n for s in sd.sparts)
        errs.append(abs(float(sd.trace().item()) - np.trace(ref).real))
        for axes in [(0,), (1,), (n - 1, 0), tuple(range(n))]:
            p = sd.probs(axes).numpy()
            _, _, p_ref, _ = oad.measure_matrix_with_probs(dims, axes, ref, orng.PRNGKey(0))
            errs.append(np.abs(p - p_ref).max())
        for keep in [(1,), (0,), (n - 1, 0)]:
            r = sd.trace_out(keep).numpy()
            errs.append(np.abs(r - oad.trace_out_matrix(dims, keep, ref, False)).max())
        for seed, axes in enumerate([(0,), (n - 1, 1), (1, 0)]):
            key = orng.PRNGKey(40 + seed)
            o, post, probs = sd.measure(axes, key)
            o_ref, post_ref, _ = oad.measure_matrix(dims, axes, ref, key)
            assert o == int(o_ref)
            got = post.gather_full().numpy()
            errs.append(np.abs(got - post_ref).max())
            assert post.local.numel() * world == got.size
            assert abs(float(post.trace().item()) - 1.0) < 1e-9
        if rank == 0:
            results.put((max(errs), sd.swaps))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dims", [(2, (4, 2, 6)), (4, (4, 4, 2, 4)), (4, (2, 6, 4, 2)), (4, (2, 6, 2, 2))])
def test_sharded_density_matrix_against_oracle(world, dims):
    err, swaps = _spawn(world, _rho_worker, dims)
    assert err < 1e-12 and swaps >= 1


# ---- whole layers with the pipelined axis swap (run_layer) ------------------------------------------


def _layer_worker(rank, world, port, kind, dims, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = len(dims)
        N = math.prod(dims)
        if kind == "vec":
            ref = ops.random_state(N, 8)
            st = ShardedStateVector.from_full(dims, torch.from_numpy(ref.reshape(-1)), rank, world, dist,
                                              NumpyBackend())
        else:
            ref = ops.random_density(N, 8)
            st = ShardedDensityMatrix.from_full(dims, torch.from_numpy(ref.reshape(-1)), rank, world, dist,
                                                NumpyBackend())
        shard_a = sorted({s.g for s in st.sparts})
        shard_b = list(range(n // 2, n // 2 + len(shard_a)))
        pipelined = 0
        for rep in range(4):
            layer = []
            for m in range(n):
                if kind == "vec":
                    layer.append((ops.random_unitary(dims[m], 10 * rep + m), (m,)))
                else:
                    layer.append((ops.loss_channel(dims[m], 0.1 + 0.05 * m), (m,)))
            for m in range(0, n - 1, 2):
                layer.append((ops.random_unitary(dims[m] * dims[m + 1], 100 + 10 * rep + m), (m, m + 1)))
            cur = set(st.sharded_axes())
            other = shard_b if cur == set(shard_a) else shard_a
            info = st.run_layer([(torch.from_numpy(o), a) for o, a in layer], next_shard=other)
            pipelined += int(info["pipelined"])
            for o, a in layer:                       # the oracle applies the PROGRAM order
                if kind == "vec":
                    ref = oad.apply_operation_vector(dims, a, ref, o, True)
                elif o.ndim == 3:
                    ref = oad.apply_kraus_matrix(dims, a, ref, o, True)
                else:
                    ref = oad.apply_operation_matrix(dims, a, ref, o, True)
            got = st.gather_full().numpy()
            err = float(np.abs(got.reshape(ref.shape) - ref).max())
            assert err < 1e-12, (rep, err)
        # a layer whose regrouping would break the program order falls back to op-by-op
        o1, o2 = ops.random_unitary(dims[0] * dims[1], 7), ops.random_unitary(dims[1], 9)
        if kind == "vec":
            st.run_layer([(torch.from_numpy(o1), (0, 1)), (torch.from_numpy(o2), (1,))], next_shard=None)
            ref = oad.apply_operation_vector(dims, (1,), oad.apply_operation_vector(dims, (0, 1), ref, o1, True), o2, True)
            assert np.abs(st.gather_full().numpy().reshape(ref.shape) - ref).max() < 1e-12
        if rank == 0:
            results.put((pipelined, st.swaps, st.packs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,world,dims", [("vec", 2, (4, 3, 2, 6)), ("vec", 4, (6, 6, 2, 6, 6)),
                                             ("rho", 2, (4, 2, 6, 2)), ("rho", 4, (4, 2, 4, 2))])
def test_run_layer_pipelined_swap_matches_program_order(kind, world, dims):
    pipelined, swaps, packs = _spawn(world, _layer_worker, kind, dims)
    assert swaps >= 3
    assert pipelined >= 2      # steady state: the outgoing axes lead, the swap hides behind S1
