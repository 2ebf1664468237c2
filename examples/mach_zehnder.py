"""
Mach-Zehnder interferometer through the container API (BASELINE.json configs[0]): a single
photon on two beam splitters with a phase shifter in one arm; the detection probability at the
first output follows sin^2(phi / 2).

    python examples/mach_zehnder.py [--points 9]
"""

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from photon_weave_b200.core import rng  # noqa: E402
from photon_weave_b200.operation import CompositeOperationType, FockOperationType, Operation  # noqa: E402
from photon_weave_b200.state.composite_envelope import CompositeEnvelope  # noqa: E402
from photon_weave_b200.state.envelope import Envelope  # noqa: E402


def interferometer(phi: float):
    e1, e2 = Envelope(), Envelope()
    e1.fock.state = 1
    ce = CompositeEnvelope(e1, e2)
    bs = Operation(CompositeOperationType.NonPolarizingBeamSplitter, eta=np.pi / 4)
    ce.apply_operation(bs, e1.fock, e2.fock)
    e1.fock.apply_operation(Operation(FockOperationType.PhaseShift, phi=phi))
    ce.apply_operation(bs, e1.fock, e2.fock)
    return ce, e1, e2


def run(points: int = 9, shots: int = 0, seed: int = 0):
    rows = []
    key = rng.PRNGKey(seed)
    for phi in np.linspace(0, np.pi, points):
        ce, e1, e2 = interferometer(float(phi))
        probs, _ = ce.measure_expectation(e1.fock)
        p = probs.cpu().numpy() if hasattr(probs, "cpu") else np.asarray(probs)
        clicks = 0
        for _ in range(shots):
            ce, e1, e2 = interferometer(float(phi))
            keys = rng.split(key, 2)
            key = keys[0]
            clicks += ce.measure(e1.fock, e2.fock, key=keys[1])[e1.fock]
        rows.append((float(phi), float(p[1]), float(np.sin(phi / 2) ** 2), clicks))
    return rows


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=9)
    ap.add_argument("--shots", type=int, default=0)
    a = ap.parse_args()
    for phi, p, want, clicks in run(a.points, a.shots):
        extra = f"  clicks {clicks}/{a.shots}" if a.shots else ""
        print(f"phi = {phi:5.3f}   P(detector 1) = {p:.12f}   sin^2(phi/2) = {want:.12f}{extra}")
