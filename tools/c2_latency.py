import sys, os, time, math
sys.path.insert(0, "/root/repo"); os.chdir("/root/repo")
import numpy as np, torch
import __graft_entry__ as g
g.build()
from oracle import operators as ops
from photon_weave_b200 import _lib
from photon_weave_b200.core import ffi
d, m = 5, 4
dims = (d,) * m
D = d ** m
gen = torch.Generator(device="cuda").manual_seed(0)
rho = torch.view_as_complex(torch.randn(D * D, 2, generator=gen, device="cuda", dtype=torch.float64))
out = torch.empty_like(rho)
bs = torch.as_tensor(np.ascontiguousarray(ops.beam_splitter(d, d, math.pi / 4))).cuda()
loss = torch.as_tensor(np.ascontiguousarray(ops.loss_channel(d, 0.1))).cuda()
u = torch.as_tensor(np.ascontiguousarray(ops.random_unitary(d, 1))).cuda()
def bench(fn, n=500):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6, e0.elapsed_time(e1) / n * 1e3
for pin in (False, True):
  if pin:
    for t_ in (bs, loss, u):
        ffi.pin_operator(t_)
  print("pinned operators (analysis cached)" if pin else "unpinned operators", flush=True)
  for thr in (1 << 18,):
    _lib.set_option("blocks_min_elems", thr)
    if True:
     for name, fn in [("bs(1,2)", lambda: ffi.apply_mat(rho, bs, dims, (1, 2), out=out)),
                     ("bs(2,3)", lambda: ffi.apply_mat(rho, bs, dims, (2, 3), out=out)),
                     ("kraus K5 (2)", lambda: ffi.kraus_mat(rho, loss, dims, (2,), out=out)),
                     ("kraus K5 (3)", lambda: ffi.kraus_mat(rho, loss, dims, (3,), out=out)),
                     ("dense 5x5 (1)", lambda: ffi.apply_mat(rho, u, dims, (1,), out=out))]:
         pass
         w, dv = bench(fn)
         print(f"thr=2^{int(math.log2(thr))} {name:14s} wall {w:7.1f} us  device {dv:7.1f} us", flush=True)
