"""Operations on a single Fock space (reference: photon_weave/operation/fock_operation.py)."""

from __future__ import annotations

from enum import Enum
from typing import Any, List, Union

import numpy as np

from .._math.ops import (
    annihilation_operator,
    creation_operator,
    displacement_operator,
    phase_operator,
    squeezing_operator,
)
from ..core import device as dev
from ..extra import interpreter
from ..state.expansion_levels import ExpansionLevel
from .helpers.fock_dimension_estimation import FockDimensions


class FockOperationType(Enum):
    """(renormalize, required_params, required_expansion_level, id) -- fock_operation.py:154-213."""

    renormalize: bool
    required_params: List[str]
    required_expansion_level: ExpansionLevel

    def __new__(cls, renormalize: bool, required_params: List[str], expansion_level: ExpansionLevel,
                op_id: int) -> "FockOperationType":
        obj = object.__new__(cls)
        obj._value_ = op_id
        obj.renormalize = renormalize
        obj.required_params = required_params
        obj.required_expansion_level = expansion_level
        return obj

    Creation = (True, [], ExpansionLevel.Vector, 1)
    Annihilation = (True, [], ExpansionLevel.Vector, 2)
    PhaseShift = (False, ["phi"], ExpansionLevel.Vector, 3)
    Squeeze = (True, ["zeta"], ExpansionLevel.Vector, 4)
    Displace = (False, ["alpha"], ExpansionLevel.Vector, 5)
    Identity = (False, [], ExpansionLevel.Vector, 6)
    Custom = (False, ["operator"], ExpansionLevel.Vector, 7)
    Expresion = (False, ["expr", "context"], ExpansionLevel.Vector, 8)

    def update(self, **kwargs: Any) -> None:
        return

    def compute_operator(self, dimensions: List[int], **kwargs: Any) -> Any:
        """fock_operation.py:218-256."""
        d = dimensions[0]
        if self is FockOperationType.Creation:
            return creation_operator(d)
        if self is FockOperationType.Annihilation:
            return annihilation_operator(d)
        if self is FockOperationType.PhaseShift:
            return phase_operator(d, kwargs["phi"])
        if self is FockOperationType.Displace:
            return displacement_operator(d, kwargs["alpha"])
        if self is FockOperationType.Squeeze:
            return squeezing_operator(d, kwargs["zeta"])
        if self is FockOperationType.Identity:
            return np.identity(d)
        if self is FockOperationType.Expresion:
            return interpreter(kwargs["expr"], kwargs["context"], dimensions)
        if self is FockOperationType.Custom:
            return kwargs["operator"]
        raise ValueError("Something went wrong in operation generation")

    def compute_dimensions(self, num_quanta: Union[int, List[int]], state: Any,
                           threshold: float = 1 - 1e-6, **kwargs: Any) -> List[int]:
        """Cut-off needed to hold the state after the operation (fock_operation.py:258-316)."""
        from .operation import Operation

        assert isinstance(num_quanta, int)
        if self in (FockOperationType.Creation, FockOperationType.Annihilation):
            return [int(num_quanta + 2)]
        if self in (FockOperationType.PhaseShift, FockOperationType.Identity):
            return [num_quanta + 1]
        if self in (FockOperationType.Displace, FockOperationType.Squeeze, FockOperationType.Expresion):
            reduced = np.array(dev.to_numpy(state), dtype=np.complex128)
            reduced = reduced / np.linalg.norm(reduced)
            cd = FockDimensions(reduced, Operation(self, **kwargs), num_quanta, threshold).compute_dimensions()
            return [max(cd, num_quanta + 1)]
        raise ValueError("Something went wrong in dimension estimation")
