"""Strawberry Fields <-> CircuitSpec bridge (reference: photon_weave/interop/strawberry_fields.py -- a stub there too)."""

from __future__ import annotations

from ..core.ir import CircuitSpec


def from_strawberry_fields(obj: object) -> CircuitSpec:
    raise NotImplementedError("Strawberry Fields interop not yet implemented; returns CircuitSpec")


def to_strawberry_fields(spec: CircuitSpec) -> object:
    raise NotImplementedError("Strawberry Fields interop not yet implemented; accepts CircuitSpec")
