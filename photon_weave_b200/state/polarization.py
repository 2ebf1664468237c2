"""
Polarization state (reference: photon_weave/state/polarization.py:34-398): one of six
labels, a 2-vector or a 2x2 density matrix in HBM, or a pointer into the Envelope /
composite product state that holds it.
"""

from __future__ import annotations

from enum import Enum
from typing import Any, Dict, Optional, Tuple, Union

import numpy as np

from ..core import device as dev
from ..photon_weave import Config
from .base_state import BaseState, check_and_renormalize
from .expansion_levels import ExpansionLevel
from .utils.measurements import (
    measure_matrix,
    measure_matrix_expectation,
    measure_matrix_jit,
    measure_vector,
    measure_vector_expectation,
    measure_vector_jit,
)
from .utils.operations import apply_operation_matrix, apply_operation_vector
from .utils.routing import route_operation
from .utils.shape_planning import build_plan
from .utils.state_transform import state_contract, state_expand

_ZERO_MSG = "The state is entirely composed of zeros, is |0⟩ attempted to be anniilated?"
_R = 1 / np.sqrt(2)


class PolarizationLabel(Enum):
    H = "H"
    V = "V"
    R = "R"
    L = "L"
    A = "A"
    D = "D"


# Jones vectors of the labels (polarization.py:139-153)
_JONES = {
    PolarizationLabel.H: (1, 0),
    PolarizationLabel.V: (0, 1),
    PolarizationLabel.R: (_R, 1j * _R),
    PolarizationLabel.L: (_R, -1j * _R),
    PolarizationLabel.A: (_R, -_R),
    PolarizationLabel.D: (_R, _R),
}


class Polarization(BaseState):
    def __init__(self, polarization: PolarizationLabel = PolarizationLabel.H,
                 envelope: Optional[Any] = None):
        super().__init__()
        self.state = polarization
        self._dimensions = 2
        self._envelope = envelope
        self._expansion_level = ExpansionLevel.Label
        self._measured = False

    @property
    def dimensions(self) -> int:
        return self._dimensions

    @dimensions.setter
    def dimensions(self, dimensions: int) -> None:
        raise ValueError("Dimensions can not be set for Polarization type, 2 by default")

    @route_operation()
    def expand(self) -> None:
        """Label -> Vector -> Matrix (polarization.py:125-164)."""
        if self.expansion_level == ExpansionLevel.Label:
            assert isinstance(self.state, PolarizationLabel)
            self.state = dev.asarray(np.array(_JONES[self.state], dtype=np.complex128)[:, None])
            self.expansion_level = ExpansionLevel.Vector
        else:
            assert dev.is_array(self.state)
            self.state, self.expansion_level = state_expand(self.state, self.expansion_level,
                                                            self.dimensions)

    @route_operation()
    def contract(self, final: ExpansionLevel = ExpansionLevel.Label, tol: float = 1e-6) -> None:
        """Matrix -> Vector when pure, Vector -> Label when it is one of the six Jones
        vectors (polarization.py:166-222)."""
        if self.measured:
            return
        if self.expansion_level is ExpansionLevel.Matrix and final < ExpansionLevel.Matrix:
            self.state, self.expansion_level, _ = state_contract(self.state, self.expansion_level)
        if self.expansion_level is ExpansionLevel.Vector and final < ExpansionLevel.Vector:
            assert dev.shape(self.state) == (self.dimensions, 1)
            vec = dev.to_numpy(self.state).reshape(2)  # two amplitudes
            for label, jones in _JONES.items():
                if np.allclose(vec, np.array(jones)):
                    self.state = label
                    self.expansion_level = ExpansionLevel.Label
                    break

    def extract(self, index: Union[int, Tuple[int, int]]) -> None:
        self.index = index
        self.state = None

    def set_index(self, minor: int, major: int = -1) -> None:
        if self.state is not None:
            raise ValueError("Polarization state does not seem to be extracted")
        self.index = (major, minor) if major >= 0 else minor

    def _set_measured(self, **kwargs: Any) -> None:
        self.measured = True
        self.state = None
        self.index = None
        self.expansion_level = None

    @route_operation()
    def measure_expectation(self) -> Tuple[Any, Any]:
        """polarization.py:259-276."""
        if self.expansion_level == ExpansionLevel.Label:
            self.expand()
        plan = build_plan([self], [self]) if Config().use_jit else None
        if self.expansion_level == ExpansionLevel.Vector:
            return measure_vector_expectation([self], [self], self.state, meta=plan)
        return measure_matrix_expectation([self], [self], self.state, meta=plan)

    @route_operation()
    def measure(self, separate_measurement: bool = False, destructive: bool = True,
                key: Any = None) -> Dict[BaseState, int]:
        """H/V basis measurement: outcome 0 = H, 1 = V (polarization.py:278-339)."""
        C = Config()
        if C.use_jit and key is None:
            raise ValueError("PRNG key is required when use_jit is enabled")
        plan = build_plan([self], [self]) if C.use_jit else None
        if self.expansion_level == ExpansionLevel.Label:
            self.expand()
        if self.expansion_level == ExpansionLevel.Vector:
            if plan is not None:
                outcomes, _, _ = measure_vector_jit([self], [self], self.state, key, meta=plan)
            else:
                outcomes, _ = measure_vector([self], [self], self.state, key=key)
        else:
            if plan is not None:
                outcomes, _, _ = measure_matrix_jit([self], [self], self.state, key, meta=plan)
            else:
                outcomes, _ = measure_matrix([self], [self], self.state, key=key)
        self.state = PolarizationLabel.H if outcomes[self] == 0 else PolarizationLabel.V
        self.expansion_level = ExpansionLevel.Label
        if destructive:
            self._set_measured()
        return outcomes

    @route_operation()
    def apply_operation(self, operation: Any) -> None:
        """polarization.py:341-398."""
        while self.expansion_level < operation.required_expansion_level:
            self.expand()
        operation.compute_dimensions(0, None)
        C = Config()
        plan = build_plan([self], [self]) if C.use_jit else None
        apply = apply_operation_vector if self.expansion_level == ExpansionLevel.Vector \
            else apply_operation_matrix
        self.state = apply([self], [self], self.state, operation.device_operator, meta=plan,
                           use_contraction=C.contractions)
        self.state = check_and_renormalize(self.state, operation.renormalize, _ZERO_MSG)
        if Config().contractions:
            self.contract()
