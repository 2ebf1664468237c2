"""Operations acting on several subsystems at once
(reference: photon_weave/operation/composite_operation.py)."""

from __future__ import annotations

from enum import Enum
from typing import Any, List, Sequence, Union

import numpy as np
from scipy.linalg import expm

from .._math import ops as _ops
from ..extra import interpreter
from ..state.expansion_levels import ExpansionLevel


def _resolve(name: Any) -> Any:
    """State classes are resolved lazily to avoid an import cycle with the state package."""
    if not isinstance(name, str):
        return name
    if name == "Fock":
        from ..state.fock import Fock

        return Fock
    if name == "Polarization":
        from ..state.polarization import Polarization

        return Polarization
    if name == "CustomState":
        from ..state.custom_state import CustomState

        return CustomState
    raise ValueError(f"Unknown state type: {name}")


class CompositeOperationType(Enum):
    """(renormalize, required_params, expected state types, expansion level, id)
    -- composite_operation.py:127-170."""

    NonPolarizingBeamSplitter = (True, ["eta"], ["Fock", "Fock"], ExpansionLevel.Vector, 1)
    CXPolarization = (True, [], ["Polarization", "Polarization"], ExpansionLevel.Vector, 2)
    SwapPolarization = (True, [], ["Polarization", "Polarization"], ExpansionLevel.Vector, 3)
    CSwapPolarization = (True, [], ["Polarization"] * 3, ExpansionLevel.Vector, 4)
    CZPolarization = (True, [], ["Polarization"] * 2, ExpansionLevel.Vector, 5)
    Expression = (True, ["expr", "state_types", "context"], [], ExpansionLevel.Vector, 6)

    renormalize: bool
    required_params: List[str]
    required_expansion_level: ExpansionLevel

    def __new__(cls, renormalize: bool, required_params: List[str], expected: Sequence[Any],
                expansion_level: ExpansionLevel, op_id: int):
        obj = object.__new__(cls)
        obj._value_ = op_id
        obj._expected = list(expected)
        obj.renormalize = renormalize
        obj.required_params = required_params
        obj.required_expansion_level = expansion_level
        return obj

    @property
    def expected_base_state_types(self) -> List[Any]:
        return [_resolve(t) for t in self._expected]

    @expected_base_state_types.setter
    def expected_base_state_types(self, value: Sequence[Any]) -> None:
        self._expected = list(value)

    def update(self, **kwargs: Any) -> None:
        """Expression operations carry their own state-type list (composite_operation.py:214-233)."""
        if self is CompositeOperationType.Expression and "state_types" in kwargs:
            self._expected = list(kwargs["state_types"])

    def compute_operator(self, dimensions: List[int], **kwargs: Any) -> Any:
        """composite_operation.py:235-299."""
        if self is CompositeOperationType.NonPolarizingBeamSplitter:
            ad_a, a = _ops.creation_operator(dimensions[0]), _ops.annihilation_operator(dimensions[0])
            ad_b, b = _ops.creation_operator(dimensions[1]), _ops.annihilation_operator(dimensions[1])
            return expm(1j * kwargs["eta"] * (np.kron(a, ad_b) + np.kron(ad_a, b)))
        if self is CompositeOperationType.CXPolarization:
            return _ops.controlled_not_operator()
        if self is CompositeOperationType.SwapPolarization:
            return _ops.swap_operator()
        if self is CompositeOperationType.CSwapPolarization:
            return _ops.controlled_swap_operator()
        if self is CompositeOperationType.CZPolarization:
            return _ops.controlled_z_operator()
        if self is CompositeOperationType.Expression:
            return interpreter(kwargs["expr"], kwargs["context"], dimensions)
        raise ValueError("Operation Type not recognized")

    def compute_dimensions(self, num_quanta: Union[int, List[int]], states: Any,
                           threshold: float = 1 - 1e6, **kwargs: Any) -> List[int]:
        """composite_operation.py:301-373."""
        from ..state.fock import Fock

        if self is CompositeOperationType.NonPolarizingBeamSplitter:
            dim = int(np.sum(np.array(num_quanta))) + 1
            return [dim, dim]
        if self in (CompositeOperationType.CXPolarization, CompositeOperationType.SwapPolarization,
                    CompositeOperationType.CZPolarization):
            return [2, 2]
        if self is CompositeOperationType.CSwapPolarization:
            return [2, 2, 2, 2]
        if self is CompositeOperationType.Expression:
            assert isinstance(num_quanta, list)
            return [num_quanta[i] + 2 if s is Fock else num_quanta[i]
                    for i, s in enumerate(self.expected_base_state_types)]
        raise ValueError("Operation Type not recognized")
