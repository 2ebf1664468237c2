"""
Envelope: one Fock space and one polarization space that travel together
(reference: photon_weave/state/envelope.py:105-1080).

While the two parts are separable each keeps its own state; ``combine()`` moves them
into one joint array (``Envelope.state``, two subsystems) in HBM; once the envelope joins
a CompositeEnvelope the array moves on into a ProductState.  All array work is a
libpwb200 kernel reached through ``core.device`` / ``core.adapters``.
"""

from __future__ import annotations

import math
import uuid
from enum import Enum
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import numpy as np

from ..constants import C0, gaussian
from ..core import device as dev
from ..core.ops import kraus_identity_check
from ..core.rng import borrow_key
from ..operation import FockOperationType, Operation, PolarizationOperationType
from ..photon_weave import Config
from .base_state import BaseState
from .composite_envelope import CompositeEnvelope
from .expansion_levels import ExpansionLevel
from .fock import Fock
from .polarization import Polarization, PolarizationLabel
from .utils import shape_planning
from .utils.measurements import (
    measure_matrix,
    measure_matrix_expectation,
    measure_matrix_jit,
    measure_POVM_matrix,
    measure_vector,
    measure_vector_expectation,
    measure_vector_jit,
)
from .utils.operations import apply_kraus_matrix, apply_operation_matrix, apply_operation_vector
from .utils.representation import representation_matrix, representation_vector
from .utils.shape_planning import ShapePlan, compiled_kernels, envelope_plan
from .utils.state_transform import state_contract, state_expand, state_expand_jit
from .utils.trace_out import trace_out_matrix, trace_out_vector


class TemporalProfile(Enum):
    Gaussian = (gaussian, {"mu": 0, "sigma": 1, "omega": None})

    def __init__(self, func: Callable, params: Any):
        self.func = func
        self.parameters = params

    def with_params(self, **kwargs: Any) -> "TemporalProfileInstance":
        params = self.parameters.copy()
        params.update(kwargs)
        return TemporalProfileInstance(self.func, params)


class TemporalProfileInstance:
    def __init__(self, func: Callable, params: Dict[Any, Any]) -> None:
        self.func = func
        self.params = params

    def get_function(self, t_a: float, omega_a: float) -> Callable:
        params = self.params.copy()
        params.update({"t_a": t_a, "omega": omega_a})
        return lambda t: self.func(t, **params)


# a 100 fs pulse (envelope.py:98-101)
default_temporal_profile = TemporalProfile.Gaussian.with_params(mu=0, sigma=42.45 * 10 ** (-15))


class Envelope:
    def __init__(self, wavelength: float = 1550, fock: Optional[Fock] = None,
                 polarization: Optional[Polarization] = None,
                 temporal_profile: TemporalProfileInstance = default_temporal_profile):
        self.uid: uuid.UUID = uuid.uuid4()
        self.fock = Fock() if fock is None else fock
        self.fock.envelope = self
        self.polarization = Polarization() if polarization is None else polarization
        self.polarization.envelope = self
        self._expansion_level: Optional[ExpansionLevel] = None
        self.composite_envelope_id: Optional[uuid.UUID] = None
        self._state: Optional[Any] = None
        self.measured = False
        self.wavelength = wavelength
        self.temporal_profile = temporal_profile
        self._plan_cache: Dict[Tuple, ShapePlan] = {}

    # -- bookkeeping --------------------------------------------------------------------
    @property
    def state(self) -> Optional[Any]:
        """The joint (fock x polarization) array in HBM, or None while the parts are separate."""
        return self._state

    @state.setter
    def state(self, value: Optional[Any]) -> None:
        self._state = dev.adopt(value)

    @property
    def expansion_level(self) -> Optional[ExpansionLevel]:
        if self._expansion_level is not None:
            return self._expansion_level
        if self.fock.expansion_level == self.polarization.expansion_level:
            return self.fock.expansion_level
        return None

    @expansion_level.setter
    def expansion_level(self, expansion_level: ExpansionLevel) -> None:
        self._expansion_level = expansion_level
        self.fock.expansion_level = expansion_level
        self.polarization.expansion_level = expansion_level

    @property
    def dimensions(self) -> int:
        return self.fock.dimensions * self.polarization.dimensions

    def __repr__(self) -> str:
        """envelope.py:162-183."""
        if self.measured:
            return "Envelope already measured"
        if self.state is None:
            left = repr(self.fock).splitlines()
            right = repr(self.polarization).splitlines()
            width = max(len(line) for line in left)
            rows = max(len(left), len(right))
            left.extend([" " * width] * (rows - len(left)))
            right.extend([""] * (rows - len(right)))
            return "\n".join(f"{a} ⊗ {b}" for a, b in zip(left, right))
        if self.expansion_level == ExpansionLevel.Vector:
            return representation_vector(self.state)
        if self.expansion_level == ExpansionLevel.Matrix:
            return representation_matrix(self.state)
        return str(self.uid)

    @property
    def composite_envelope(self) -> Optional[CompositeEnvelope]:
        if self.composite_envelope_id is not None:
            ce = CompositeEnvelope._instances[self.composite_envelope_id][0]
            assert ce is not None, "Composite Envelope should exist"
            return ce
        return None

    def set_composite_envelope_id(self, uid: uuid.UUID) -> None:
        self.composite_envelope_id = uid

    def _current_state_order(self) -> List[BaseState]:
        """Tensor order of the two subsystems inside ``self.state`` (envelope.py:937-949)."""
        if isinstance(self.fock.index, int) and isinstance(self.polarization.index, int):
            ordered: List[Optional[BaseState]] = [None, None]
            ordered[self.fock.index] = self.fock
            ordered[self.polarization.index] = self.polarization
            return [s for s in ordered if s is not None]
        return [self.fock, self.polarization]

    def _plan_for_states(self, targets: Tuple[BaseState, ...]) -> ShapePlan:
        order = tuple(self._current_state_order())
        key = (tuple(s.dimensions for s in order), tuple(id(s) for s in order),
               tuple(id(t) for t in targets))
        if key not in self._plan_cache:
            self._plan_cache[key] = shape_planning.build_plan(order, targets)
        return self._plan_cache[key]

    def prepare_meta(self, *target_states: BaseState):
        return envelope_plan(self, *target_states).meta

    def compiled_kernels(self, *target_states: BaseState):
        return compiled_kernels(envelope_plan(self, *target_states))

    # -- representation changes ---------------------------------------------------------
    def combine(self) -> None:
        """Fock (x) Polarization -> one joint array (envelope.py:185-235)."""
        for s in (self.fock, self.polarization):
            if s.measured:
                raise ValueError("Parts of this envelope have already been destructively measured,"
                                 " cannot combine")
        if self.fock.expansion_level == ExpansionLevel.Label:
            self.fock.expand()
        if self.polarization.expansion_level == ExpansionLevel.Label:
            self.polarization.expand()
        while self.fock.expansion_level < self.polarization.expansion_level:
            self.fock.expand()
        while self.polarization.expansion_level < self.fock.expansion_level:
            self.polarization.expand()
        level = self.fock.expansion_level
        assert level == self.polarization.expansion_level
        if level in (ExpansionLevel.Vector, ExpansionLevel.Matrix):
            self.state = dev.kron(self.fock.state, self.polarization.state)
            self.fock.extract(0)
            self.polarization.extract(1)
            self.expansion_level = level

    def expand(self, *args: Any, **kwargs: Any) -> None:
        """envelope.py:250-279 (the extra arguments exist for the routing decorator)."""
        if self.state is None:
            self.fock.expand()
            self.polarization.expand()
        else:
            assert isinstance(self.expansion_level, ExpansionLevel)
            expand = state_expand_jit if Config().use_jit else state_expand
            self.state, self.expansion_level = expand(self.state, self.expansion_level,
                                                      self.dimensions)

    def contract(self, final: ExpansionLevel = ExpansionLevel.Vector, tol: float = 1e-6) -> None:
        """envelope.py:830-853."""
        assert dev.is_array(self.state)
        if self.expansion_level == ExpansionLevel.Vector:
            return
        assert dev.shape(self.state) == (self.dimensions, self.dimensions)
        if self.expansion_level == ExpansionLevel.Matrix:
            self.state, self.expansion_level, _ = state_contract(self.state, self.expansion_level)

    def reorder(self, *states: BaseState) -> None:
        """Swap the two subsystems when they are not in the requested order
        (envelope.py:750-828)."""
        order = list(states)
        if len(order) == 2:
            if type(order[0]) is type(order[1]) and isinstance(order[0], (Fock, Polarization)):
                raise ValueError("Given states have to be unique")
        elif len(order) > 2:
            raise ValueError("Too many states given")
        for s in order:
            if s is not self.polarization and s is not self.fock:
                raise ValueError("Given states have to be members of the envelope, "
                                 "use env.fock and env.polarization")
        if self.state is None:
            return
        if len(order) == 1:
            order.append(self.polarization if order[0] is self.fock else self.fock)
        current = self._current_state_order()
        if current[0] is order[0] and current[1] is order[1]:
            return
        dims = [s.dimensions for s in current]
        is_matrix = self.expansion_level == ExpansionLevel.Matrix
        if is_matrix:
            assert dev.shape(self.state) == (self.dimensions, self.dimensions)
        else:
            assert dev.shape(self.state) == (self.dimensions, 1)
        self.state = dev.permute(self.state, dims, [1, 0], is_matrix)
        self.fock.index, self.polarization.index = self.polarization.index, self.fock.index

    def resize_fock(self, new_dimensions: int, state: Optional[Fock] = None) -> bool:
        """Pad or (when no weight is lost) truncate the Fock axis (envelope.py:995-1077)."""
        if self.state is None:
            return self.fock.resize(new_dimensions)
        is_matrix = self.expansion_level == ExpansionLevel.Matrix
        shrink = new_dimensions < self.fock.dimensions or (is_matrix and new_dimensions == self.fock.dimensions)
        if shrink:
            # the reference looks at the reduced Fock state here, and its trace_out leaves the
            # envelope in (fock, polarization) order (envelope.py:1034, 923)
            self.reorder(self.fock, self.polarization)
        order = self._current_state_order()
        dims = [s.dimensions for s in order]
        axis = int(self.fock.index)
        if shrink:
            levels, _, _ = dev.occupancy(self.state, dims, is_matrix)
            if levels[axis] >= new_dimensions:
                return False
        elif new_dimensions == self.fock.dimensions:
            return False
        self.state = dev.resize(self.state, dims, axis, new_dimensions, is_matrix)
        self.fock.dimensions = new_dimensions
        return True

    def _set_measured(self, remove_composite: bool = True) -> None:
        if self.composite_envelope is not None and remove_composite:
            self.composite_envelope.envelopes.remove(self)
            self.composite_envelope_id = None
        self.measured = True
        self.state = None

    # -- measurement --------------------------------------------------------------------
    def measure(self, *states: BaseState, separate_measurement: bool = False,
                destructive: bool = True, key: Any = None, return_key: bool = False):
        """envelope.py:281-438."""
        if self.measured:
            raise ValueError("Envelope has already been destroyed")
        for s in states:
            assert s in [self.fock, self.polarization]
        if len(states) == 0:
            states = (self.fock, self.polarization)
        C = Config()
        if C.use_jit and key is None:
            raise ValueError("PRNG key is required when use_jit is enabled")
        if key is None:
            key = C.random_key
        outcomes: Dict[BaseState, int] = {}
        next_key = key
        if self.state is None:
            for s in (self.polarization, self.fock):
                use_key, key = borrow_key(key)
                out = s.measure(separate_measurement=separate_measurement, destructive=destructive,
                                key=use_key)
                outcomes.update(out)
            next_key = key
        else:
            self.reorder(self.fock, self.polarization)
            if not separate_measurement and len(states) == 1:
                states = (self.fock, self.polarization)
            if len(states) == 2:
                separate_measurement = True
            assert all(isinstance(s.index, int) for s in states)
            pair = [self.fock, self.polarization]
            is_vector = self.expansion_level == ExpansionLevel.Vector
            if C.use_jit:
                plan = self._plan_for_states(tuple(states))
                use_key, _ = borrow_key(key)
                fn = measure_vector_jit if is_vector else measure_matrix_jit
                outcomes, self.state, next_key = fn(pair, list(states), self.state, use_key, meta=plan)
            else:
                fn = measure_vector if is_vector else measure_matrix
                outcomes, self.state = fn(pair, list(states), self.state, key=key)
            for s in pair:
                if separate_measurement and s not in states:
                    s.state = self.state
                    s.index = None
                    s.expansion_level = self.expansion_level
                    s.contract()
                    if C.contractions:
                        s.contract()
                else:
                    s.state = dev.basis_vector(s.dimensions, outcomes[s])
                    s.expansion_level = ExpansionLevel.Vector
                    s.index = None
                    if destructive:
                        s._set_measured()
                    else:
                        while s.expansion_level > ExpansionLevel.Label and C.contractions:
                            s.contract()
        if destructive:
            self._set_measured()
        if return_key:
            return outcomes, next_key
        return outcomes

    def measure_expectation(self, *states: BaseState) -> Tuple[Any, Any]:
        """envelope.py:440-475."""
        if self.measured:
            raise ValueError("Envelope has already been destroyed")
        if len(states) == 0:
            states = (self.fock, self.polarization)
        if self.state is None:
            self.combine()
        plan = self._plan_for_states(tuple(states))
        fn = measure_vector_expectation if self.expansion_level == ExpansionLevel.Vector \
            else measure_matrix_expectation
        return fn(self._current_state_order(), list(states), self.state, meta=plan)

    def measure_POVM(self, operators: List[Any], *states: BaseState, destructive: bool = True,
                     key: Any = None) -> Tuple[int, Dict[BaseState, int]]:
        """envelope.py:477-583."""
        if self.measured:
            raise ValueError("This envelope was already measured")
        if len(states) == 2:
            if type(states[0]) is type(states[1]) and isinstance(states[0], (Fock, Polarization)):
                raise ValueError("Given states have to be unique")
        elif len(states) > 2:
            raise ValueError("Too many states given")
        for s in states:
            if s is not self.polarization and s is not self.fock:
                raise ValueError("Given states have to be members of the envelope, "
                                 "use env.fock and env.polarization")
        if len(states) == 1 and self.state is None:
            return states[0].measure_POVM(operators, destructive=destructive, key=key)
        if self.composite_envelope_id is not None:
            return self.composite_envelope.measure_POVM(operators, *states,
                                                        destructive=destructive, key=key)
        while self.expansion_level < ExpansionLevel.Matrix:
            self.expand()
        if len(states) == 2 and self.state is None:
            self.combine()
        assert isinstance(self.fock.index, int) and isinstance(self.polarization.index, int)
        self.reorder(self.fock, self.polarization)
        plan = self._plan_for_states(tuple(states))
        pair = [self.fock, self.polarization]
        outcome, self.state = measure_POVM_matrix(pair, list(states), operators, self.state, key,
                                                  meta=plan.meta)
        outcome = int(outcome)
        if len(states) == 2:
            if destructive:
                self._set_measured()
                self.fock._set_measured()
                self.polarization._set_measured()
            return (outcome, {})
        if destructive:
            other = self.fock if states[0] is self.polarization else self.polarization
            other.state = trace_out_matrix(pair, [other], self.state)
            other.index = None
            other.expansion_level = ExpansionLevel.Matrix
            other.contract()
            states[0]._set_measured()
            self._set_measured()
        return (outcome, {})

    # -- channels and operations --------------------------------------------------------
    def apply_kraus(self, operators: List[Any], *states: BaseState, validity_check: bool = True,
                    key: Any = None) -> None:
        """envelope.py:585-669."""
        if len(states) == 0:
            raise ValueError("At least one state must be defined")
        uids = [s.uid for s in states]
        if len(states) == 2 and (self.fock.uid not in uids or self.polarization.uid not in uids):
            raise ValueError("Both given states must belong to the Envelope")
        if len(states) > 2:
            raise ValueError("Too many states given")
        if self.state is None and any(isinstance(s.index, tuple) for s in states):
            assert self.composite_envelope_id is not None
            self.composite_envelope.apply_kraus(operators, *states)
            return
        if len(states) == 1 and self.state is None:
            states[0].apply_kraus(operators)
            return
        if self.state is None:
            self.combine()
        if validity_check and not kraus_identity_check(operators):
            raise ValueError("Kraus operators do not sum to the identity sum K^dagg K != I")
        self.reorder(*states)
        while self.expansion_level < ExpansionLevel.Matrix:
            self.expand()
        self.state = apply_kraus_matrix(self._current_state_order(), states, self.state, operators,
                                        use_contraction=Config().contractions)
        self.contract()
        if Config().contractions:
            self.contract()

    def trace_out(self, *states: BaseState) -> Any:
        """Reduced state of the given subsystems, in the given order (envelope.py:881-935)."""
        if self.composite_envelope is not None:
            return self.composite_envelope.trace_out(*states)
        if self.state is None:
            if len(states) == 1:
                assert isinstance(states[0], (Polarization, Fock))
                return states[0].state
            if len(states) == 2:
                self.combine()
        self.reorder(self.fock, self.polarization)
        assert self.state is not None
        pair = [self.fock, self.polarization]
        if self.expansion_level == ExpansionLevel.Vector:
            return trace_out_vector(pair, states, self.state)
        return trace_out_matrix(pair, states, self.state)

    def apply_operation(self, operation: Operation, *states: Union[Fock, Polarization]) -> None:
        """envelope.py:1079-1080 and above: route to the part while the envelope is separable,
        otherwise act on the joint array."""
        kind = operation._operation_type
        if isinstance(kind, FockOperationType) and not isinstance(states[0], Fock):
            raise ValueError(f"Cannot apply {type(kind)} to {type(states[0])}")
        if isinstance(kind, PolarizationOperationType) and not isinstance(states[0], Polarization):
            raise ValueError(f"Cannot apply {type(kind)} to {type(states[0])}")
        if self.state is None:
            states[0].apply_operation(operation)
            return
        _, biggest = dev.norm_stats(self.state)
        if not biggest > 0:
            raise ValueError("The state is invalid.The state only consists of 0s.")
        C = Config()
        if C.use_jit and C.dynamic_dimensions:
            raise ValueError("JIT mode does not support dynamic dimension resizing; "
                             "pre-size states or disable use_jit.")
        if isinstance(kind, FockOperationType) and isinstance(states[0], Fock):
            operation.compute_dimensions(states[0]._num_quanta, self.fock.trace_out())
            target_dim = operation.dimensions[0]
            if C.use_jit and target_dim != states[0].dimensions:
                raise ValueError("JIT mode requires static Fock dimensions; "
                                 f"expected {states[0].dimensions}, got {target_dim}. "
                                 "Pre-size the state or disable use_jit.")
            if not C.use_jit:
                states[0].resize(target_dim)
        elif isinstance(kind, PolarizationOperationType) and isinstance(states[0], Polarization):
            operation.compute_dimensions([0], None)
        order = self._current_state_order()
        plan = self._plan_for_states(tuple(states)) if C.use_jit else None
        apply = apply_operation_vector if self.expansion_level == ExpansionLevel.Vector \
            else apply_operation_matrix
        self.state = apply(order, list(states), self.state, operation.device_operator, meta=plan,
                           use_contraction=C.contractions)
        total, biggest = dev.norm_stats(self.state)
        if not biggest > 0:
            raise ValueError("The state is entirely composed of zeros, "
                             "is |0⟩ attempted to be annihilated?")
        if operation.renormalize:
            self.state = dev.divide(self.state, math.sqrt(total))
        if Config().contractions:
            self.contract()

    def overlap_integral(self, other: "Envelope", delay: float, n: float = 1) -> float:
        """Overlap of the temporal profiles, ``other`` delayed by ``delay`` seconds
        (envelope.py:951-977)."""
        from scipy.integrate import quad

        f1 = self.temporal_profile.get_function(t_a=0, omega_a=(C0 / n) / self.wavelength)
        f2 = other.temporal_profile.get_function(t_a=delay, omega_a=(C0 / n) / other.wavelength)
        result, _ = quad(lambda x: np.conj(f1(x)) * f2(x), -np.inf, np.inf)
        return result
