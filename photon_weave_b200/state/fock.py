"""
Fock state (reference: photon_weave/state/fock.py:35-506): a number-basis label, a state
vector or a density matrix in HBM, or a pointer (``index``) into the Envelope /
composite product state that holds it.
"""

from __future__ import annotations

from typing import Any, Dict, Optional, Tuple, Union

import numpy as np

from ..core import device as dev
from ..core.ops import num_quanta_matrix, num_quanta_vector
from ..photon_weave import Config
from .base_state import BaseState, check_and_renormalize
from .expansion_levels import ExpansionLevel
from .utils.measurements import (
    measure_matrix,
    measure_matrix_expectation,
    measure_matrix_jit,
    measure_pnr_matrix,
    measure_pnr_vector,
    measure_vector,
    measure_vector_expectation,
    measure_vector_jit,
)
from .utils.operations import apply_operation_matrix, apply_operation_vector
from .utils.routing import route_operation
from .utils.shape_planning import build_plan
from .utils.state_transform import state_contract, state_expand

_ZERO_MSG = "The state is entirely composed of zeros, is |0⟩ attempted to be annihilated?"


class Fock(BaseState):
    def __init__(self, envelope: Optional[Any] = None):
        super().__init__()
        self.state = 0
        self.envelope = envelope
        self.expansion_level = ExpansionLevel.Label
        self.measured = False

    def __hash__(self) -> int:
        return hash(self.uid)

    def __eq__(self, other: Any) -> bool:
        """Equal labels, or equally shaped arrays that are numerically close (fock.py:77-103)."""
        if not isinstance(other, Fock):
            return False
        if isinstance(self.state, int) and isinstance(other.state, int):
            return self.state == other.state
        if dev.is_array(self.state) and dev.is_array(other.state):
            if dev.shape(self.state) == dev.shape(other.state):
                return bool(np.allclose(dev.to_numpy(self.state), dev.to_numpy(other.state)))
        return False

    @route_operation()
    def expand(self) -> None:
        """Label -> Vector -> Matrix (fock.py:105-124).  An unsized label gets state + 3 levels."""
        assert self.state is not None and self.expansion_level is not None
        if self.dimensions < 0 and isinstance(self.state, int):
            self.dimensions = self.state + 3
        self.state, self.expansion_level = state_expand(self.state, self.expansion_level, self.dimensions)

    def contract(self, final: ExpansionLevel = ExpansionLevel.Label, tol: float = 1e-6) -> None:
        """fock.py:126-151."""
        if self.measured:
            return
        assert self.state is not None and self.expansion_level is not None
        success = True
        while self.expansion_level > final and success:
            self.state, self.expansion_level, success = state_contract(self.state, self.expansion_level)

    def extract(self, index: Union[int, Tuple[int, int]]) -> None:
        """The state moved into a product space; keep only the pointer (fock.py:153-161)."""
        self.index = index
        self.state = None

    @property
    def _num_quanta(self) -> int:
        """Highest number state with non-zero weight, -1 on failure (fock.py:163-190)."""
        if isinstance(self.state, int):
            return self.state
        arr = self.state if dev.is_array(self.state) else (self.trace_out() if self.state is None else None)
        if arr is not None and dev.is_array(arr):
            if dev.shape(arr) == (self.dimensions, 1):
                return int(num_quanta_vector(arr))
            if dev.shape(arr) == (self.dimensions, self.dimensions):
                return int(num_quanta_matrix(arr))
        return -1

    def set_index(self, minor: int, major: int = -1) -> None:
        self.index = (major, minor) if major >= 0 else minor

    @route_operation()
    def measure(self, separate_measurement: bool = False, destructive: bool = True,
                key: Any = None) -> Dict[BaseState, int]:
        """Number-basis measurement of a free-standing Fock state (fock.py:211-283)."""
        C = Config()
        if C.use_jit and key is None:
            raise ValueError("PRNG key is required when use_jit is enabled")
        plan = build_plan([self], [self]) if C.use_jit else None
        outcomes: Dict[BaseState, int]
        if self.expansion_level == ExpansionLevel.Label:
            assert isinstance(self.state, int)
            outcomes = {self: self.state}
        elif self.expansion_level == ExpansionLevel.Vector:
            if plan is not None:
                outcomes, _, _ = measure_vector_jit([self], [self], self.state, key, meta=plan)
            else:
                outcomes, _ = measure_vector([self], [self], self.state, key=key)
        else:
            if plan is not None:
                outcomes, _, _ = measure_matrix_jit([self], [self], self.state, key, meta=plan)
            else:
                outcomes, _ = measure_matrix([self], [self], self.state, key=key)
        self.state = outcomes[self]
        self.expansion_level = ExpansionLevel.Label
        if destructive:
            self._set_measured()
        if self.envelope is not None and not separate_measurement:
            if not self.envelope.polarization.measured:
                out = self.envelope.polarization.measure(separate_measurement=separate_measurement,
                                                         destructive=destructive)
                outcomes.update(out)
        return outcomes

    def _set_measured(self, **kwargs: Any) -> None:
        self.measured = True
        self.state = None
        self.index = None
        self.expansion_level = None

    @route_operation()
    def measure_expectation(self) -> Tuple[Any, Any]:
        """fock.py:295-309."""
        if self.expansion_level == ExpansionLevel.Label:
            self.expand()
        plan = build_plan([self], [self]) if Config().use_jit else None
        if self.expansion_level == ExpansionLevel.Vector:
            return measure_vector_expectation([self], [self], self.state, meta=plan)
        return measure_matrix_expectation([self], [self], self.state, meta=plan)

    @route_operation()
    def measure_pnr(self, dark_rate: float = 0.0, efficiency: float = 1.0, detection_window: float = 1.0,
                    jitter_std: float = 0.0, destructive: bool = True, key: Any = None):
        """Photon-number-resolving detection with loss, dark counts and jitter (fock.py:311-372)."""
        if self.expansion_level == ExpansionLevel.Label:
            self.expand()
        C = Config()
        if C.use_jit and key is None:
            raise ValueError("PRNG key is required when use_jit is enabled")
        plan = build_plan([self], [self]) if C.use_jit else None
        fn = measure_pnr_vector if self.expansion_level == ExpansionLevel.Vector else measure_pnr_matrix
        outcomes, _, jitter, next_key = fn([self], [self], self.state, efficiency=efficiency,
                                           dark_rate=dark_rate, detection_window=detection_window,
                                           jitter_std=jitter_std, key=key, meta=plan)
        if next_key is None:
            next_key = key
        self.state = outcomes[self]
        self.expansion_level = ExpansionLevel.Label
        if destructive:
            self._set_measured()
        return outcomes, jitter, next_key

    @route_operation()
    def resize(self, new_dimensions: int) -> bool:
        """Pad, or truncate when no weight is lost (fock.py:374-437)."""
        if new_dimensions < 1:
            return False
        if self.index is not None:
            return False
        if self.expansion_level is ExpansionLevel.Label:
            assert isinstance(self.state, int)
            if self.state > new_dimensions - 1:
                return False
            self.dimensions = new_dimensions
            return False  # the reference falls through to ``return False`` here (fock.py:396-401, 437)
        is_matrix = self.expansion_level is ExpansionLevel.Matrix
        if new_dimensions > self.dimensions:
            self.state = dev.resize(self.state, (self.dimensions,), 0, new_dimensions, is_matrix)
            self.dimensions = new_dimensions
            return True
        if new_dimensions < self.dimensions:
            if is_matrix:
                # fock.py:430-434 truncates a density matrix without looking at its support
                pass
            elif not num_quanta_vector(self.state) < new_dimensions + 1:
                return False
            self.state = dev.resize(self.state, (self.dimensions,), 0, new_dimensions, is_matrix)
            self.dimensions = new_dimensions
            return True
        return False

    @route_operation()
    def apply_operation(self, operation: Any) -> None:
        """fock.py:439-499."""
        while self.expansion_level < operation.required_expansion_level:
            self.expand()
        C = Config()
        if C.use_jit and C.dynamic_dimensions:
            raise ValueError("Dynamic dimension resizing is not supported when use_jit=True; "
                             "precompute cutoffs before compiling.")
        if C.dynamic_dimensions:
            operation.compute_dimensions(self._num_quanta, self.trace_out())
            self.resize(operation.dimensions[0])
        else:
            operation._dimensions = [self.dimensions]
        plan = build_plan([self], [self]) if C.use_jit else None
        apply = apply_operation_vector if self.expansion_level == ExpansionLevel.Vector else apply_operation_matrix
        self.state = apply([self], [self], self.state, operation.device_operator, meta=plan,
                           use_contraction=C.contractions)
        self.state = check_and_renormalize(self.state, operation.renormalize, _ZERO_MSG)
        if Config().contractions:
            self.contract()
