"""Operations on a CustomState (reference: photon_weave/operation/custom_state_operation.py)."""

from __future__ import annotations

from enum import Enum
from typing import Any, List, Union

from ..extra import interpreter
from ..state.expansion_levels import ExpansionLevel


class CustomStateOperationType(Enum):
    Expresion = (True, ["expr", "context"], ExpansionLevel.Vector, 1)
    Custom = (True, ["operator"], ExpansionLevel.Vector, 2)

    renormalize: bool
    required_params: List[str]
    required_expansion_level: ExpansionLevel

    def __new__(cls, renormalize: bool, required_params: List[str], expansion_level: ExpansionLevel,
                op_id: int) -> "CustomStateOperationType":
        obj = object.__new__(cls)
        obj._value_ = op_id
        obj.renormalize = renormalize
        obj.required_params = required_params
        obj.required_expansion_level = expansion_level
        return obj

    def update(self, **kwargs: Any) -> None:
        return

    def compute_operator(self, dimensions: List[int], **kwargs: Any) -> Any:
        if self is CustomStateOperationType.Custom:
            return kwargs["operator"]
        if self is CustomStateOperationType.Expresion:
            return interpreter(kwargs["expr"], kwargs["context"], dimensions)
        raise ValueError("Operator not recognized")

    def compute_dimensions(self, num_quanta: Union[int, List[int]], state: Any, threshold: float = 1,
                           **kwargs: Any) -> List[int]:
        return [int(state.shape[0])]
