import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from oracle import operators as ops
from photon_weave_b200.core import ffi
dims = (3, 4, 5, 2)
n = int(np.prod(dims))
jobs = []
for j in range(7):
    psi = ops.random_state(n, 100 + j)
    oplist = [(0, (j % 4,), ops.random_unitary(dims[j % 4], j)), (0, (3, 1), ops.random_unitary(8, 50 + j))]
    if j % 2:
        oplist.append((0, (2,), ops.random_unitary(5, 70 + j)))
    jobs.append((psi, oplist))
want = [ffi.circuit_host(psi, False, dims, oplist) for psi, oplist in jobs]
want2 = [ffi.circuit_host(psi, False, dims, oplist) for psi, oplist in jobs]
print("serial deterministic:", [bool(np.array_equal(a, b)) for a, b in zip(want, want2)])
for nbuf in (2, 3, 4):
    for mode in ("all", "one"):
        ins = [torch.as_tensor(psi).pin_memory().numpy() for psi, _ in jobs]
        outs = [torch.zeros(n, dtype=torch.complex128).pin_memory().numpy() for _ in jobs]
        with ffi.HostPipeline(dims, False, np.complex128, buffers=nbuf) as pl:
            if mode == "all":
                tickets = [pl.submit(ins[j], outs[j], jobs[j][1]) for j in range(len(jobs))]
                for tk in reversed(tickets):
                    pl.wait(tk)
            else:
                for j in range(len(jobs)):
                    pl.wait(pl.submit(ins[j], outs[j], jobs[j][1]))
        print(nbuf, mode, [float(np.abs(outs[j].reshape(-1) - want[j].reshape(-1)).max()) for j in range(len(jobs))])
