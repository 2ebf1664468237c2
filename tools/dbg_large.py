import sys, time, faulthandler
faulthandler.enable()
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from oracle import adapters_np as oad
from oracle import operators as oops
from photon_weave_b200.photon_weave import Config
from photon_weave_b200.operation import CompositeOperationType, FockOperationType, Operation
from photon_weave_b200.state.composite_envelope import CompositeEnvelope
from photon_weave_b200.state.envelope import Envelope
def P(*a):
    torch.cuda.synchronize(); print(time.strftime("%X"), *a, flush=True)
Config().set_contraction(False)
n, d = 8, 6
dims = (d,) * n
envs = [Envelope() for _ in range(n)]
for e in envs: e.fock.dimensions = d
focks = [e.fock for e in envs]
ce = CompositeEnvelope(*envs)
P("combine")
ce.combine(*focks)
ps = ce.product_states[0]
psi = oops.random_state(d**n, 3)
ps.state = psi.reshape(-1, 1)
P("apply displace")
ce.apply_operation(Operation(FockOperationType.Displace, alpha=0.3), focks[5])
P("apply bs")
ce.apply_operation(Operation(CompositeOperationType.NonPolarizingBeamSplitter, eta=np.pi / 4), focks[6], focks[1])
P("apply squeeze")
ce.apply_operation(Operation(FockOperationType.Squeeze, zeta=0.1 + 0.1j), focks[7])
P("trace")
red = ce.trace_out(focks[7], focks[2])
P("reorder")
ce.reorder(*focks)
got = ps.state.cpu().numpy()
P("expect")
probs, rest = ce.measure_expectation(*focks[:7])
P("done", probs.shape, rest.shape)
