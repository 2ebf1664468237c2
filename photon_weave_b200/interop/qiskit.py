"""Qiskit <-> CircuitSpec bridge (reference: photon_weave/interop/qiskit.py -- a stub there too)."""

from __future__ import annotations

from ..core.ir import CircuitSpec


def from_qiskit(obj: object) -> CircuitSpec:
    raise NotImplementedError("Qiskit interop not yet implemented; returns CircuitSpec")


def to_qiskit(spec: CircuitSpec) -> object:
    raise NotImplementedError("Qiskit interop not yet implemented; accepts CircuitSpec")
