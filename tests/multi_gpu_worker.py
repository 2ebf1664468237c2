"""
Worker of tests/test_gpu_multi.py: launched by torchrun with one rank per GPU (NCCL), compares
the sharded CUDA paths (peer-memory swaps and NCCL swaps) with the NumPy oracle on every rank's
reassembled result.  Prints MULTI_OK on rank 0 when everything matched.
"""

import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

from oracle import adapters_np as oad  # noqa: E402  (the checker)
from oracle import operators as ops  # noqa: E402
from oracle import rng_threefry as orng  # noqa: E402
from photon_weave_b200.sharded import (  # noqa: E402
    CudaBackend, PeerBuffers, ShardedDensityMatrix, ShardedStateVector)

TOL = 1e-12


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


def make_state(cls, dims, full, rank, world, local_rank, use_peers):
    gd = dims if cls is ShardedStateVector else tuple(dims) + tuple(dims)
    vaxes, sparts = cls.initial_layout(dims, world)
    own = cls._own_block(gd, torch.as_tensor(full).reshape(-1), sparts, rank).cuda()
    peers, spare = None, None
    if use_peers:
        peers = PeerBuffers(own.numel(), torch.complex128, 2, rank, world, dist, local_rank)
        peers.tensors[0].copy_(own)
        own, spare = peers.tensors[0], peers.tensors[1]
    return cls(dims, own, vaxes, sparts, rank, world, dist, CudaBackend(), spare=spare, peers=peers), peers


def check_vector(rank, world, local_rank, use_peers):
    dims = (6, 6, 4, 6, 6, 2)
    N = math.prod(dims)
    psi = ops.random_state(N, 31)
    st, peers = make_state(ShardedStateVector, dims, psi, rank, world, local_rank, use_peers)
    ref = psi
    program = [((1,), ops.random_unitary(6, 1)), ((0,), ops.random_unitary(6, 2)),          # t <= 8 on a sharded axis
               ((3, 4), ops.beam_splitter(6, 6, 0.7)), ((0, 1), ops.beam_splitter(6, 6, 0.4)),  # t = 36, block-structured
               ((2,), ops.phase_operator(4, 0.3)), ((4, 0), ops.random_unitary(36, 3)),          # t = 36 dense on a sharded axis
               ((5, 3), ops.random_unitary(12, 4)), ((1, 0), ops.beam_splitter(6, 6, 1.1))]
    for axes, op in program:
        st.apply_operation(dev(op), axes)
        ref = oad.apply_operation_vector(dims, axes, ref, op, True)
        err = rel(st.gather_full().cpu().numpy().reshape(-1, 1), ref)
        assert err < TOL, ("vector apply", axes, err, use_peers)
    assert st.swaps >= 2
    for axes in [(0,), (5, 0), (1, 2, 3)]:
        p = st.probs(axes).cpu().numpy()
        _, _, p_ref, _ = oad.measure_vector_with_probs(dims, axes, ref, orng.PRNGKey(0))
        assert np.abs(p - p_ref).max() < TOL, ("vector probs", axes)
    for keep in [(1,), (0, 5)]:
        r = st.reduced(keep).cpu().numpy()
        assert rel(r, oad.trace_out_vector(dims, keep, ref, False)) < TOL, ("reduced", keep)
    for seed, axes in enumerate([(0,), (5, 1), (2, 0)]):
        key = orng.PRNGKey(70 + seed)
        o, post, _ = st.measure(axes, key)
        o_ref, post_ref, _ = oad.measure_vector(dims, axes, ref, key)
        assert o == int(o_ref), ("vector outcome", axes, o, int(o_ref))
        assert rel(post.gather_full().cpu().numpy(), post_ref.reshape(-1)) < 1e-11, ("vector post", axes)
    swaps = (st.swaps, st.fused_swaps)
    del st
    torch.cuda.synchronize()
    dist.barrier()
    if peers is not None:
        peers.close()
    return swaps


def check_density(rank, world, local_rank, use_peers):
    dims = (8, 4, 8, 2)
    N = math.prod(dims)
    rho = ops.random_density(N, 17)
    sd, peers = make_state(ShardedDensityMatrix, dims, rho, rank, world, local_rank, use_peers)
    ref = rho
    program = [("k", (1,), ops.loss_channel(4, 0.2)), ("k", (0,), ops.loss_channel(8, 0.1)),
               ("u", (0, 2), ops.beam_splitter(8, 8, 0.6)), ("u", (2,), ops.random_unitary(8, 5)),
               ("k", (2,), np.stack([ops.random_unitary(8, 6 + k) for k in range(3)]) / math.sqrt(3.0)),
               ("u", (1, 0), ops.random_unitary(32, 9)), ("k", (0,), ops.loss_channel(8, 0.3))]
    for kind, axes, op in program:
        if kind == "u":
            sd.apply_operation(dev(op), axes)
            ref = oad.apply_operation_matrix(dims, axes, ref, op, True)
        else:
            sd.apply_kraus(dev(op), axes)
            ref = oad.apply_kraus_matrix(dims, axes, ref, op, True)
        err = rel(sd.gather_full().cpu().numpy(), ref)
        assert err < TOL, ("rho apply", kind, axes, err, use_peers)
        assert all(s.g < len(dims) for s in sd.sparts)
    assert sd.swaps >= 2
    assert abs(float(sd.trace().item()) - np.trace(ref).real) < 1e-12
    for axes in [(0,), (2, 0), (0, 1, 2)]:
        p = sd.probs(axes).cpu().numpy()
        _, _, p_ref, _ = oad.measure_matrix_with_probs(dims, axes, ref, orng.PRNGKey(0))
        assert np.abs(p - p_ref).max() < TOL, ("rho probs", axes)
    for keep in [(1,), (0,), (2, 0)]:
        r = sd.trace_out(keep).cpu().numpy()
        assert rel(r, oad.trace_out_matrix(dims, keep, ref, False)) < TOL, ("trace_out", keep)
    for seed, axes in enumerate([(0,), (2, 1)]):
        key = orng.PRNGKey(90 + seed)
        o, post, _ = sd.measure(axes, key)
        o_ref, post_ref, _ = oad.measure_matrix(dims, axes, ref, key)
        assert o == int(o_ref), ("rho outcome", axes)
        assert rel(post.gather_full().cpu().numpy(), post_ref) < 1e-11, ("rho post", axes)
    swaps = sd.swaps
    del sd
    torch.cuda.synchronize()
    dist.barrier()
    if peers is not None:
        peers.close()
    return swaps


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__ as g

    g.build()
    out = {}
    for use_peers in (True, False):
        out[("peer" if use_peers else "nccl") + "_vector_swaps"] = check_vector(rank, world, local_rank, use_peers)
        out[("peer" if use_peers else "nccl") + "_rho_swaps"] = check_density(rank, world, local_rank, use_peers)
    dist.barrier()
    if rank == 0:
        print("MULTI_OK", out, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
