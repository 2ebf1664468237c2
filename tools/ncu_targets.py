#!/usr/bin/env python
"""The launches `ncu --set full` captures for profiles/r2_traffic.json: every kernel class of the
final dispatch three times on the full-size inputs (configs[4] vector, configs[3] density matrix).
    python tools/ncu_targets.py            # plain run (must exit 0 before the ncu run)
"""
import math
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402

g.build()
from oracle import operators as ops  # noqa: E402
from photon_weave_b200.core import ffi  # noqa: E402


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


REPS = 3
HEADLINE = len(sys.argv) > 1 and sys.argv[1] == "headline"   # only the three classes bench.py quotes
LATE = len(sys.argv) > 1 and sys.argv[1] == "late"           # the kernels added after r2_ncu_all_classes.json


def late_round():
    """complex64 fiber-pair kernels, complex64 TF32 x 3 and complex128 Gauss two-sided kernels."""
    gen = torch.Generator(device="cuda").manual_seed(0)
    d, modes = 6, 12
    dims = (d,) * modes
    N = d**modes
    a = torch.view_as_complex(torch.randn(N, 2, generator=gen, device="cuda", dtype=torch.float32))
    b = torch.empty_like(a)
    c64 = lambda x: dev(x).to(torch.complex64)
    u, bs = c64(ops.random_unitary(d, 1)), c64(ops.beam_splitter(d, d, math.pi / 4))
    for _ in range(2):
        ffi.apply_vec(a, u, dims, (0,), out=b)          # k_apply_small_c64x2
        ffi.apply_vec(a, bs, dims, (0, 1), out=b)       # k_apply_blocks_c64x2
        ffi.reduced_from_vec(a, dims, (3,))             # k_reduced_small<.., PAIR>
    torch.cuda.synchronize()
    del a, b
    torch.cuda.empty_cache()
    d, modes = 8, 5
    dims = (d,) * modes
    D = d**modes
    for dt in (torch.float32, torch.float64):
        rho = torch.view_as_complex(torch.randn(D * D, 2, generator=gen, device="cuda", dtype=dt))
        out = torch.empty_like(rho)
        cast = (lambda x: dev(x).to(torch.complex64)) if dt == torch.float32 else dev
        loss, u8 = cast(ops.loss_channel(d, 0.1)), cast(ops.random_unitary(d, 1))
        bs8 = cast(ops.beam_splitter(d, d, math.pi / 4))
        for _ in range(2):
            if dt == torch.float32:
                ffi.kraus_mat(rho, loss, dims, (2,), out=out)     # k_apply_blocks_c64x2 (superoperator)
            ffi.apply_mat(rho, u8, dims, (2,), out=out)           # k_two_sided_mma (TF32 x 3 / Gauss DMMA)
            ffi.apply_mat(rho, bs8, dims, (1, 2), out=out)        # k_two_sided_mma, super-blocks
            ffi.apply_mat(rho, bs8, dims, (3, 4), out=out)        # k_two_sided_mma_rows
            if dt == torch.float64:
                ffi.kraus_mat(rho, loss, dims, (4,), out=out)     # k_apply_blocks_seg_mma
        torch.cuda.synchronize()
        del rho, out
        torch.cuda.empty_cache()
    print("ncu targets (late) done")


if LATE:
    late_round()
    sys.exit(0)
d, modes = 6, 12
dims = (d,) * modes
N = d**modes
gen = torch.Generator(device="cuda").manual_seed(0)
a = torch.view_as_complex(torch.randn(N, 2, generator=gen, device="cuda", dtype=torch.float64))
b = torch.empty_like(a)
u, ph = dev(ops.random_unitary(d, 1)), dev(ops.phase_operator(d, 0.3))
bs, dense2 = dev(ops.beam_splitter(d, d, math.pi / 4)), dev(ops.random_unitary(d * d, 2))
for _ in range(REPS):
    ffi.apply_vec(a, u, dims, (0,), out=b)          # k_apply_small
    ffi.apply_vec(a, bs, dims, (0, 1), out=b)       # k_apply_blocks
    if HEADLINE:
        continue
    ffi.apply_vec(a, u, dims, (9,), out=b)          # k_apply_small, short inner run
    ffi.apply_vec(a, u, dims, (11,), out=b)         # k_apply_contig_mma
    ffi.apply_vec(a, ph, dims, (9,), out=b)         # k_apply_diag
    ffi.apply_vec(a, bs, dims, (10, 11), out=b)     # k_apply_blocks_seg_mma
    ffi.apply_vec(a, dense2, dims, (3, 4), out=b)   # k_apply_dense_mma
    ffi.apply_vec_multi(a, [u, u, u], dims, (3, 4, 5), out=b)   # k_apply_multi
torch.cuda.synchronize()
del a, b
torch.cuda.empty_cache()
d, modes = 8, 5
dims = (d,) * modes
D = d**modes
rho = torch.view_as_complex(torch.randn(D * D, 2, generator=gen, device="cuda", dtype=torch.float64))
out = torch.empty_like(rho)
loss, u8 = dev(ops.loss_channel(d, 0.1)), dev(ops.random_unitary(d, 1))
bs8 = dev(ops.beam_splitter(d, d, math.pi / 4))
for _ in range(REPS):
    ffi.kraus_mat(rho, loss, dims, (2,), out=out)         # k_apply_blocks on the superoperator
    if HEADLINE:
        continue
    ffi.kraus_mat(rho, loss, dims, (4,), out=out)         # k_apply_blocks_seg_mma
    ffi.apply_mat(rho, u8, dims, (2,), out=out)           # k_two_sided_mma
    ffi.apply_mat(rho, bs8, dims, (1, 2), out=out)        # k_two_sided_mma, super-blocks
    ffi.apply_mat(rho, bs8, dims, (3, 4), out=out)        # k_two_sided_mma_rows
torch.cuda.synchronize()
print("ncu targets done")
