"""Piquasso <-> CircuitSpec bridge (reference: photon_weave/interop/piquasso.py -- a stub there too)."""

from __future__ import annotations

from ..core.ir import CircuitSpec


def from_piquasso(obj: object) -> CircuitSpec:
    raise NotImplementedError("Piquasso interop not yet implemented; returns CircuitSpec")


def to_piquasso(spec: CircuitSpec) -> object:
    raise NotImplementedError("Piquasso interop not yet implemented; accepts CircuitSpec")
