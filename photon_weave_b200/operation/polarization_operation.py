"""Single-qubit gates on a polarization space
(reference: photon_weave/operation/polarization_operation.py)."""

from __future__ import annotations

from enum import Enum
from typing import Any, List, Union

from .._math import ops as _ops
from ..state.expansion_levels import ExpansionLevel


class PolarizationOperationType(Enum):
    renormalize: bool
    required_params: List[str]
    required_expansion_level: ExpansionLevel

    def __new__(cls, renormalize: bool, required_params: List[str], expansion_level: ExpansionLevel,
                op_id: int) -> "PolarizationOperationType":
        obj = object.__new__(cls)
        obj._value_ = op_id
        obj.renormalize = renormalize
        obj.required_params = required_params
        obj.required_expansion_level = expansion_level
        return obj

    I = (True, [], ExpansionLevel.Vector, 1)  # noqa: E741
    X = (True, [], ExpansionLevel.Vector, 2)
    Y = (True, [], ExpansionLevel.Vector, 3)
    Z = (True, [], ExpansionLevel.Vector, 4)
    H = (True, [], ExpansionLevel.Vector, 5)
    S = (True, [], ExpansionLevel.Vector, 6)
    T = (True, [], ExpansionLevel.Vector, 7)
    SX = (True, [], ExpansionLevel.Vector, 8)
    RX = (True, ["theta"], ExpansionLevel.Vector, 9)
    RY = (True, ["theta"], ExpansionLevel.Vector, 10)
    RZ = (True, ["theta"], ExpansionLevel.Vector, 11)
    U3 = (True, ["phi", "theta", "omega"], ExpansionLevel.Vector, 12)
    Custom = (True, ["operator"], ExpansionLevel.Vector, 13)

    def update(self, **kwargs: Any) -> None:
        return

    def compute_operator(self, dimensions: List[int], **kwargs: Any) -> Any:
        """polarization_operation.py:262-325."""
        fixed = {
            PolarizationOperationType.I: _ops.identity_operator,
            PolarizationOperationType.X: _ops.x_operator,
            PolarizationOperationType.Y: _ops.y_operator,
            PolarizationOperationType.Z: _ops.z_operator,
            PolarizationOperationType.H: _ops.hadamard_operator,
            PolarizationOperationType.S: _ops.s_operator,
            PolarizationOperationType.T: _ops.t_operator,
            PolarizationOperationType.SX: _ops.sx_operator,
        }
        if self in fixed:
            return fixed[self]()
        if self is PolarizationOperationType.RX:
            return _ops.rx_operator(kwargs["theta"])
        if self is PolarizationOperationType.RY:
            return _ops.ry_operator(kwargs["theta"])
        if self is PolarizationOperationType.RZ:
            return _ops.rz_operator(kwargs["theta"])
        if self is PolarizationOperationType.U3:
            return _ops.u3_operator(kwargs["phi"], kwargs["theta"], kwargs["omega"])
        if self is PolarizationOperationType.Custom:
            return kwargs["operator"]
        raise ValueError("Operator not recognized")

    def compute_dimensions(self, num_quanta: Union[int, List[int]], state: Any, threshold: float = 1,
                           **kwargs: Any) -> List[int]:
        return [2]
