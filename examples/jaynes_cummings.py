"""
Jaynes-Cummings time evolution through the container API (BASELINE.json configs[2]:
Fock cutoff 64 (x) qubit): one photon in the cavity, the qubit in its lower basis state,
H = w a^dagger a - (w/2) sigma_z + g (a sigma_+ + a^dagger sigma_-) with g = w = pi.

The same program as the reference's examples/jaynes_cummings_model.py, written against
photon_weave_b200 (only the import root differs), plus the analytic check: the initial state
couples to |2>|1> with strength sqrt(2) g at detuning 2 w, so the population of qubit level 1
is p(t) = (2/3) sin^2(sqrt(3) pi t).

    python examples/jaynes_cummings.py [--cutoff 64] [--steps 300]
"""

import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from photon_weave_b200.operation import CompositeOperationType, Operation  # noqa: E402
from photon_weave_b200.photon_weave import Config  # noqa: E402
from photon_weave_b200.state.composite_envelope import CompositeEnvelope  # noqa: E402
from photon_weave_b200.state.custom_state import CustomState  # noqa: E402
from photon_weave_b200.state.envelope import Envelope  # noqa: E402
from photon_weave_b200.state.fock import Fock  # noqa: E402


def run(cutoff: int = 64, steps: int = 300, t_delta: float = 0.01):
    cfg = Config()
    cfg.set_contraction(True)
    cfg.set_dynamic_dimensions(False)
    env = Envelope()
    env.fock.state = 1
    env.fock.dimensions = cutoff
    qubit = CustomState(2)
    ce = CompositeEnvelope(env, qubit)
    ce.combine(env.fock, qubit)

    w = g = np.pi
    lower = lambda n: np.diag(np.sqrt(np.arange(1, n)), 1)  # noqa: E731
    ctx = {
        "a": lambda d: lower(d[0]), "a_dag": lambda d: lower(d[0]).T, "i_p": lambda d: np.eye(d[0]),
        "s_z": lambda d: np.diag([1.0, -1.0]), "s_plus": lambda d: np.array([[0, 1], [0, 0]]),
        "s_minus": lambda d: np.array([[0, 0], [1, 0]]), "s_i": lambda d: np.eye(2),
    }
    hamiltonian = ("add",
                   ("s_mult", w, ("kron", ("m_mult", "a_dag", "a"), "s_i")),
                   ("s_mult", -w / 2, ("kron", "i_p", "s_z")),
                   ("s_mult", g, ("add", ("kron", "a", "s_plus"), ("kron", "a_dag", "s_minus"))))
    step = Operation(CompositeOperationType.Expression, expr=("expm", ("s_mult", -1j, hamiltonian, t_delta)),
                     context=ctx, state_types=[Fock, CustomState])
    populations = []
    t0 = time.perf_counter()
    for _ in range(steps):
        ce.apply_operation(step, env.fock, qubit)
        reduced = qubit.trace_out()
        reduced = reduced.cpu().numpy() if hasattr(reduced, "cpu") else np.asarray(reduced)
        populations.append(float(np.real(reduced[1, 1])))
    per_step = (time.perf_counter() - t0) / steps
    t = (np.arange(steps) + 1) * t_delta
    analytic = (2.0 / 3.0) * np.sin(np.sqrt(3.0) * np.pi * t) ** 2
    return np.array(populations), analytic, per_step, env.fock.dimensions


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cutoff", type=int, default=64)
    ap.add_argument("--steps", type=int, default=300)
    a = ap.parse_args()
    pops, analytic, per_step, dims = run(a.cutoff, a.steps)
    print(f"steps {a.steps}  cutoff {a.cutoff} (now {dims})  {per_step * 1e6:.1f} us/step  "
          f"max |p - analytic| = {np.abs(pops - analytic).max():.2e}")
