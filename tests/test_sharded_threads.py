"""Randomised multi-rank cases of the sharded state vector / density matrix (layout planning, full
and group-wise axis swaps, pack passes, pipelined layers, reductions, measurement) against the
oracle, on the in-process rank harness of tests/thread_dist.py: world sizes 2..8 over mixed
dimensions, many more shapes than the gloo tests can afford."""
import math

import numpy as np
import pytest
import torch

from oracle import adapters_np as oad
from oracle import operators as ops
from oracle import rng_threefry as orng
from photon_weave_b200.sharded import ShardedDensityMatrix, ShardedStateVector
from test_sharded_gloo import NumpyBackend
from thread_dist import run_ranks


def _random_dims(rng, world, n_lo, n_hi, limit, choices=(2, 3, 4, 5, 6, 8)):
    """Dimensions whose leading axes can absorb the world size, total size <= limit."""
    for _ in range(200):
        n = int(rng.integers(n_lo, n_hi + 1))
        dims = tuple(int(rng.choice(choices)) for _ in range(n))
        if math.prod(dims) > limit:
            continue
        try:
            ShardedStateVector.shard_plan(dims, world)
        except ValueError:
            continue
        return dims
    return None


def _vec_program(rank, world, dist, dims, seed):
    rng = np.random.default_rng(seed)           # the same stream on every rank
    N, n = math.prod(dims), len(dims)
    psi = ops.random_state(N, seed)
    st = ShardedStateVector.from_full(dims, torch.from_numpy(psi.reshape(-1)), rank, world, dist, NumpyBackend())
    ref = psi
    worst = 0.0
    for step in range(6):
        k = int(rng.integers(1, min(2, n) + 1))
        axes = tuple(int(a) for a in rng.choice(n, size=k, replace=False))
        t = math.prod(dims[a] for a in axes)
        op = ops.random_unitary(t, 100 * seed + step)
        try:
            st.apply_operation(torch.from_numpy(op), axes)
        except RuntimeError as exc:             # a layout that cannot supply the factors: must say so
            assert "no admissible axis swap" in str(exc)
            return None
        ref = oad.apply_operation_vector(dims, axes, ref, op, True)
        worst = max(worst, float(np.abs(st.gather_full().numpy().reshape(-1, 1) - ref).max()))
    axes = tuple(int(a) for a in rng.choice(n, size=min(2, n), replace=False))
    p = st.probs(axes).numpy()
    _, _, p_ref, _ = oad.measure_vector_with_probs(dims, axes, ref, orng.PRNGKey(1))
    worst = max(worst, float(np.abs(p - p_ref).max()))
    key = orng.PRNGKey(seed)
    try:
        o, post, _ = st.measure(axes, key)
    except ValueError as exc:      # raised on every rank BEFORE any collective: nothing hangs
        assert "cannot be spread over" in str(exc)
        return worst
    o_ref, post_ref, _ = oad.measure_vector(dims, axes, ref, key)
    assert o == int(o_ref)
    if post_ref.size >= world:                 # the post state is spread over all ranks again
        worst = max(worst, float(np.abs(post.gather_full().numpy().reshape(-1, 1) - post_ref).max()))
    return worst


@pytest.mark.parametrize("world", [2, 3, 4, 6, 8])
def test_random_vector_programs(world):
    rng = np.random.default_rng(1000 + world)
    ran = 0
    for case in range(12):
        dims = _random_dims(rng, world, 3, 5, 3000)
        if dims is None:
            continue
        res = run_ranks(world, _vec_program, dims, 7 * case + world)
        if res[0] is None:
            assert all(r is None for r in res)     # every rank agreed that the swap is impossible
            continue
        assert max(res) < 1e-12, (dims, res)
        ran += 1
    assert ran >= 3


def _layer_program(rank, world, dist, kind, dims, seed):
    n, N = len(dims), math.prod(dims)
    cls = ShardedStateVector if kind == "vec" else ShardedDensityMatrix
    ref = ops.random_state(N, seed) if kind == "vec" else ops.random_density(N, seed)
    st = cls.from_full(dims, torch.from_numpy(ref.reshape(-1)), rank, world, dist, NumpyBackend())
    rng = np.random.default_rng(seed)
    shard_a = sorted({s.g for s in st.sparts})
    pipelined = 0
    for rep in range(3):
        order = [int(m) for m in rng.permutation(n)]
        layer = []
        for m in order:
            if kind == "vec":
                layer.append((ops.random_unitary(dims[m], 10 * rep + m), (m,)))
            else:
                layer.append((ops.loss_channel(dims[m], 0.1 + 0.03 * m), (m,)))
        start = int(rng.integers(0, 2))
        for m in range(start, n - 1, 2):
            layer.append((ops.random_unitary(dims[m] * dims[m + 1], 50 + 10 * rep + m), (m, m + 1)))
        cur = set(st.sharded_axes())
        others = [g for g in range(n) if g not in cur]
        nxt = shard_a if cur != set(shard_a) else others[:max(1, len(shard_a))]
        try:
            info = st.run_layer([(torch.from_numpy(o), a) for o, a in layer], next_shard=nxt)
        except RuntimeError as exc:
            assert "no admissible axis swap" in str(exc)
            return None
        pipelined += int(info["pipelined"])
        for o, a in layer:
            if kind == "vec":
                ref = oad.apply_operation_vector(dims, a, ref, o, True)
            elif o.ndim == 3:
                ref = oad.apply_kraus_matrix(dims, a, ref, o, True)
            else:
                ref = oad.apply_operation_matrix(dims, a, ref, o, True)
        err = float(np.abs(st.gather_full().numpy().reshape(ref.shape) - ref).max())
        assert err < 1e-12, (dims, rep, err)
        if kind == "rho":
            assert all(s.g < n for s in st.sparts)   # only row subsystems are ever sharded
    return pipelined


@pytest.mark.parametrize("kind,world,limit", [("vec", 2, 3000), ("vec", 4, 3000), ("vec", 8, 5000),
                                              ("rho", 2, 64), ("rho", 4, 64), ("rho", 8, 96)])
def test_random_layers_through_run_layer(kind, world, limit):
    rng = np.random.default_rng(2000 + world + (0 if kind == "vec" else 50))
    ran, piped = 0, 0
    for case in range(8):
        # a layer needs local axes that can take over the whole world size while its two-mode
        # operators are applied: many small even dimensions (the bench shapes are 6^12 and 8^5)
        dims = _random_dims(rng, world, 4, 5 if kind == "rho" else 6, limit, choices=(2, 2, 4, 6))
        if dims is None:
            continue
        res = run_ranks(world, _layer_program, kind, dims, 11 * case + world)
        if res[0] is None:
            assert all(r is None for r in res)
            continue
        assert len(set(res)) == 1                  # every rank took the same schedule
        ran += 1
        piped += res[0]
    assert ran >= 3


def _rho_program(rank, world, dist, dims, seed):
    rng = np.random.default_rng(seed)
    n, N = len(dims), math.prod(dims)
    rho = ops.random_density(N, seed)
    sd = ShardedDensityMatrix.from_full(dims, torch.from_numpy(rho.reshape(-1)), rank, world, dist, NumpyBackend())
    ref, worst = rho, 0.0
    for step in range(4):
        k = int(rng.integers(1, min(2, n) + 1))
        axes = tuple(int(a) for a in rng.choice(n, size=k, replace=False))
        t = math.prod(dims[a] for a in axes)
        try:
            if step % 2 == 0:
                op = ops.random_unitary(t, 10 * seed + step)
                sd.apply_operation(torch.from_numpy(op), axes)
                ref = oad.apply_operation_matrix(dims, axes, ref, op, True)
            else:
                ks = ops.loss_channel(t, 0.2)
                sd.apply_kraus(torch.from_numpy(ks), axes)
                ref = oad.apply_kraus_matrix(dims, axes, ref, ks, True)
        except RuntimeError as exc:
            assert "no admissible axis swap" in str(exc)
            return None
        assert all(s.g < n for s in sd.sparts)
        worst = max(worst, float(np.abs(sd.gather_full().numpy() - ref).max()))
    worst = max(worst, abs(float(sd.trace().item()) - float(np.trace(ref).real)))
    axes = tuple(int(a) for a in rng.choice(n, size=min(2, n), replace=False))
    p = sd.probs(axes).numpy()
    _, _, p_ref, _ = oad.measure_matrix_with_probs(dims, axes, ref, orng.PRNGKey(0))
    worst = max(worst, float(np.abs(p - p_ref).max()))
    keep = tuple(int(a) for a in rng.choice(n, size=1, replace=False))
    try:
        r = sd.trace_out(keep).numpy()
        worst = max(worst, float(np.abs(r - oad.trace_out_matrix(dims, keep, ref, False)).max()))
    except RuntimeError as exc:
        assert "no admissible axis swap" in str(exc)
    key = orng.PRNGKey(seed + 5)
    try:
        o, post, _ = sd.measure(axes, key)
    except ValueError as exc:
        assert "cannot be spread over" in str(exc)
        return worst
    o_ref, post_ref, _ = oad.measure_matrix(dims, axes, ref, key)
    assert o == int(o_ref)
    worst = max(worst, float(np.abs(post.gather_full().numpy() - post_ref).max()))
    return worst


@pytest.mark.parametrize("world", [2, 4, 8])
def test_random_density_matrix_programs(world):
    rng = np.random.default_rng(3000 + world)
    ran = 0
    for case in range(16):
        dims = _random_dims(rng, world, 3, 5, 72 if world < 8 else 200, choices=(2, 2, 3, 4, 6, 8))
        if dims is None:
            continue
        res = run_ranks(world, _rho_program, dims, 17 * case + world)
        if res[0] is None:
            assert all(r is None for r in res)
            continue
        assert max(res) < 1e-12, (dims, res)
        ran += 1
    assert ran >= 3


# ---- the schedule bench.py --gpus N runs (layouts, LayerRunner, pipelined swap) on CPU ranks ----


def _bench_schedule(rank, world, dist, kind, modes, cutoff):
    import bench
    from photon_weave_b200 import bench_sharded as bs

    dims = (cutoff,) * modes
    cls = ShardedStateVector if kind == "vec" else ShardedDensityMatrix
    layer = bench.make_layer(modes, cutoff) if kind == "vec" else bench.make_rho_layer(modes, cutoff)
    vaxes, sparts, shard_a, shard_b = bs._layouts(cls, dims, world, modes, layer)
    ref = bench.preflight_input(kind, dims)
    local = cls._own_block(dims if kind == "vec" else dims + dims, torch.from_numpy(ref.reshape(-1)), sparts, rank)
    st = cls(dims, local, vaxes, sparts, rank, world, dist, NumpyBackend())
    d_ops = [torch.from_numpy(np.ascontiguousarray(o)) for _, o, _ in layer]
    runner = bs.LayerRunner(st, layer, d_ops, shard_a, shard_b, bs._apply_vec if kind == "vec" else bs._apply_rho,
                            "pipelined")
    piped = 0
    for _ in range(3):
        runner.run()
        piped += int(runner.last_info["pipelined"])
        for t, o, _ in layer:                      # the oracle applies the PROGRAM order
            ref = bench.oracle_apply(kind, dims, t, ref, o)
        err = float(np.abs(st.gather_full().numpy().reshape(ref.shape) - ref).max())
        assert err < 1e-12, err
    return piped, st.swaps, st.packs, tuple(shard_a), tuple(shard_b)


@pytest.mark.parametrize("kind,world,modes,cutoff", [("vec", 2, 6, 6), ("vec", 4, 6, 6), ("vec", 8, 8, 4),
                                                    ("rho", 2, 4, 4), ("rho", 4, 4, 4), ("rho", 8, 5, 4)])
def test_bench_schedule_is_pipelined_and_exact(kind, world, modes, cutoff):
    res = run_ranks(world, _bench_schedule, kind, modes, cutoff)
    assert len(set(res)) == 1
    piped, swaps, packs, shard_a, shard_b = res[0]
    assert swaps == 3 and not set(shard_a) & set(shard_b)
    assert piped >= 2          # steady state of the two alternating layouts: every swap is hidden


def _bad_arguments(rank, world, dist):
    dims = (4, 3, 2)
    psi = ops.random_state(24, 1)
    st = ShardedStateVector.from_full(dims, torch.from_numpy(psi.reshape(-1)), rank, world, dist, NumpyBackend())
    u3 = torch.from_numpy(ops.random_unitary(3, 1))
    for bad_axes in [(), (3,), (-1,), (1, 1)]:
        with pytest.raises(ValueError):
            st.apply_operation(u3, bad_axes)
    with pytest.raises(AssertionError):
        st.apply_operation(u3, (0,))                       # 3 x 3 operator on a 4-level subsystem
    with pytest.raises(AssertionError):
        st.run_layer([(torch.from_numpy(ops.loss_channel(3, 0.1)), (1,))])   # Kraus stack on a vector
    st.apply_operation(u3.to(torch.complex64), (1,))       # promoted to the state's dtype
    ref = oad.apply_operation_vector(dims, (1,), psi, u3.numpy().astype(np.complex64).astype(np.complex128), True)
    err = float(np.abs(st.gather_full().numpy().reshape(-1, 1) - ref).max())
    rho = ops.random_density(24, 2)
    sd = ShardedDensityMatrix.from_full(dims, torch.from_numpy(rho.reshape(-1)), rank, world, dist, NumpyBackend())
    with pytest.raises(ValueError):
        sd.apply_kraus(torch.from_numpy(ops.loss_channel(3, 0.1)), (3,))     # column subsystems are not addressable
    with pytest.raises(AssertionError):
        sd.apply_operation(u3, (2,))
    return err


def test_sharded_argument_errors_follow_the_seam():
    assert max(run_ranks(2, _bad_arguments)) < 1e-6
