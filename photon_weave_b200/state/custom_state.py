"""
CustomState: a d-level system with no built-in physics (reference:
photon_weave/state/custom_state.py:37-402).  It is never held by an Envelope, so its
index is ``None`` or a ``(product_state, position)`` tuple.
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np

from ..core import adapters
from ..core import device as dev
from ..core.ops import apply_kraus, kraus_identity_check
from ..core.rng import borrow_key
from ..operation.custom_state_operation import CustomStateOperationType
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

_ZERO_MSG = "The state is entirely composed of zeros, is |0⟩ attempted to be annihilated?"


class CustomState(BaseState):
    __slots__ = ("_dimensions_set",)

    def __init__(self, dimensions: int):
        super().__init__()
        self._dimensions_set = False
        self.dimensions = dimensions
        self.state = 0
        self.expansion_level = ExpansionLevel.Label

    @property
    def dimensions(self) -> int:
        return self._dimensions

    @dimensions.setter
    def dimensions(self, dimensions: int) -> None:
        if self._dimensions_set:
            raise ValueError("Dimensions can only be set once")
        self._dimensions = dimensions
        self._dimensions_set = True

    @property
    def index(self) -> Union[None, Tuple[int, int]]:
        assert not isinstance(self._index, int)
        return self._index

    @index.setter
    def index(self, index: Union[None, Tuple[int, int]]) -> None:
        assert not isinstance(index, int)
        self._index = index

    # a CustomState survives measurement (custom_state.py:141-149)
    @property
    def measured(self) -> bool:
        return False

    @measured.setter
    def measured(self, measured: bool) -> None:
        pass

    def _set_measured(self, **kwargs: Any) -> None:
        pass

    @route_operation()
    def expand(self) -> None:
        assert self.state is not None and isinstance(self.expansion_level, ExpansionLevel)
        self.state, self.expansion_level = state_expand(self.state, self.expansion_level,
                                                        self.dimensions)

    @route_operation()
    def contract(self, final: ExpansionLevel = ExpansionLevel.Label, tol: float = 1e-6) -> None:
        assert isinstance(self.expansion_level, ExpansionLevel)
        success = True
        while self.expansion_level > final and success:
            self.state, self.expansion_level, success = state_contract(self.state,
                                                                       self.expansion_level)

    def extract(self, index: Union[int, Tuple[int, int]]) -> None:
        assert isinstance(index, tuple) and len(index) == 2
        self.index = index
        self.state = None

    def set_index(self, minor: int = -1, major: int = -1) -> None:
        if minor == -1 and major == -1:
            self.index = None
        elif major >= 0 and minor >= 0:
            self.index = (major, minor)
        else:
            raise ValueError("Either set both parameters (minor, major) or none of them")

    @route_operation()
    def measure(self, separate_measurement: bool = False, destructive: bool = True,
                key: Any = None) -> Dict[BaseState, int]:
        """Basis measurement; the state collapses to the label (custom_state.py:180-243)."""
        C = Config()
        if C.use_jit and key is None:
            raise ValueError("PRNG key is required when use_jit is enabled")
        plan = build_plan([self], [self]) if C.use_jit else None
        if self.expansion_level == ExpansionLevel.Label:
            assert isinstance(self.state, int)
            return {self: self.state}
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
        self.state = outcomes[self]
        self.expansion_level = ExpansionLevel.Label
        return outcomes

    @route_operation()
    def measure_expectation(self) -> Tuple[Any, Any]:
        if self.expansion_level == ExpansionLevel.Label:
            self.expand()
        plan = build_plan([self], [self]) if Config().use_jit else None
        if self.expansion_level == ExpansionLevel.Vector:
            return measure_vector_expectation([self], [self], self.state, meta=plan)
        return measure_matrix_expectation([self], [self], self.state, meta=plan)

    def measure_POVM(self, operators: List[Any], destructive: bool = True, partial: bool = False,
                     key: Any = None) -> Tuple[int, Dict[BaseState, int]]:
        """custom_state.py:262-306: p_k = Tr(E_k rho) (not Tr(E_k rho E_k^dagger) as in the
        other containers), post = E rho E^dagger / Tr."""
        if self.index is not None:
            if self.composite_envelope is None or not hasattr(self.composite_envelope, "measure_POVM"):
                raise ValueError("Composite envelope not attached for measurement")
            return self.composite_envelope.measure_POVM(operators, self)
        for op in operators:
            assert tuple(op.shape) == (self.dimensions, self.dimensions)
        while self.expansion_level < ExpansionLevel.Matrix:
            self.expand()
        # operator-sized host arithmetic on the d x d reduced state decides the branch ...
        rho = dev.to_numpy(self.state)
        host_ops = [np.asarray(dev.to_numpy(op), dtype=np.complex128) for op in operators]
        probabilities = np.array([np.trace(op @ rho).real for op in host_ops])
        probabilities = probabilities / probabilities.sum()
        if key is None:
            key = Config().random_key
        use_key, _ = borrow_key(key)
        outcome = dev.draw(probabilities, use_key)
        # ... and the sampled branch runs through the two-sided kernel
        post = adapters.apply_operation_matrix((self.dimensions,), (0,), self.state,
                                               dev.asarray(host_ops[outcome]))
        _, trace, _ = dev.occupancy(post, (self.dimensions,), True)
        self.state = dev.divide(post, trace)
        if Config().contractions:
            self.contract()
        return (outcome, {})

    def apply_kraus(self, operators: List[Any], identity_check: bool = True) -> None:
        """custom_state.py:308-345."""
        if self.index is not None:
            self.composite_envelope.apply_kraus(operators, self)
            return
        while self.expansion_level != ExpansionLevel.Matrix:
            self.expand()
        for op in operators:
            if tuple(op.shape) != (self.dimensions, self.dimensions):
                raise ValueError(f"Kraus operator has incorrect dimensions: {tuple(op.shape)}, "
                                 f"expected ({self.dimensions},{self.dimensions})")
        if identity_check and not kraus_identity_check(operators):
            raise ValueError("Invalid Kraus Channel")
        self.state = apply_kraus(self.state, operators)
        if Config().contractions:
            self.contract()

    @route_operation()
    def apply_operation(self, operation: Any) -> None:
        """custom_state.py:347-402."""
        assert isinstance(operation._operation_type, CustomStateOperationType)
        while self.expansion_level < operation.required_expansion_level:
            self.expand()
        operation.compute_dimensions(0, self.trace_out())
        assert tuple(operation.operator.shape) == (self.dimensions, self.dimensions)
        C = Config()
        plan = build_plan([self], [self]) if C.use_jit else None
        apply = apply_operation_vector if self.expansion_level == ExpansionLevel.Vector \
            else apply_operation_matrix
        self.state = apply([self], [self], self.state, operation.device_operator, meta=plan,
                           use_contraction=C.contractions)
        self.state = check_and_renormalize(self.state, operation.renormalize, _ZERO_MSG)
        if Config().contractions:
            self.contract()
