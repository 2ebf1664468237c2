"""
Serializable circuit description (reference: photon_weave/core/ir.py:16-110): plain data,
no dependency on the state containers.  ``run_circuit`` executes a spec on the CUDA seam.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Iterable, Optional, Tuple


@dataclass(frozen=True)
class StateSpec:
    dims: Tuple[int, ...]
    rep: str  # "vector" | "matrix"
    data: Optional[Any] = None


@dataclass(frozen=True)
class OpSpec:
    name: str
    operator: Any
    targets: Tuple[int, ...]
    rep: str = "vector"
    use_contraction: Optional[bool] = None


@dataclass(frozen=True)
class MeasureSpec:
    targets: Tuple[int, ...]
    rep: str = "vector"
    povm: Optional[Any] = None


@dataclass(frozen=True)
class TraceOutSpec:
    targets: Tuple[int, ...]


@dataclass(frozen=True)
class CircuitSpec:
    state: StateSpec
    steps: Tuple[object, ...]

    def with_steps(self, steps: Iterable[object]) -> "CircuitSpec":
        return CircuitSpec(self.state, tuple(steps))
