"""
CompositeEnvelope / ProductState: many envelopes (and CustomStates) sharing product
spaces (reference: photon_weave/state/composite_envelope.py:64-1457).

B200-first differences from the reference, none of them observable through the API:

* **Logical vs physical order.**  The reference physically transposes the whole product
  state every time the tensor order changes (``reorder`` before every Kraus channel, POVM,
  multi-state trace-out and -- through ``resize_fock`` -- once per Fock space after EVERY
  ``apply_operation``).  The kernels here address subsystems through strides, so a
  ``ProductState`` keeps the array in whatever *physical* order it was built in and
  ``reorder`` only permutes the *logical* list ``state_objs``.  The array is permuted (one
  ``pw_permute`` pass) only if somebody reads ``ProductState.state`` while the two orders
  differ.
* **Fused bookkeeping.**  After an operation the reference runs two any-nonzero passes, a
  norm, a division and -- per Fock space -- a full ``psi psi^dagger`` partial trace to find
  the highest occupied level.  Here ONE ``pw_occupancy`` pass returns the highest occupied
  level of every subsystem together with the norm and the all-zero test;
  the pre-check is skipped when the array is the one verified by the previous call.
"""

from __future__ import annotations

import math
import uuid
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np

from ..core import adapters
from ..core import device as dev
from ..core.ops import kraus_identity_check, kron_reduce
from ..core.rng import borrow_key
from ..operation import (
    CompositeOperationType,
    CustomStateOperationType,
    FockOperationType,
    Operation,
    PolarizationOperationType,
)
from ..photon_weave import Config
from .expansion_levels import ExpansionLevel
from .utils import shape_planning
from .utils.measurements import (
    measure_matrix_expectation,
    measure_matrix_jit,
    measure_POVM_matrix,
    measure_vector_expectation,
    measure_vector_jit,
)
from .utils.operations import apply_kraus_matrix, apply_operation_matrix, apply_operation_vector
from .utils.shape_planning import ShapePlan, compiled_kernels
from .utils.state_transform import state_contract, state_expand, state_expand_jit

BaseState = Any
Envelope = Any


def _is_envelope(obj: object) -> bool:
    return hasattr(obj, "fock") and hasattr(obj, "polarization")


def _is_custom_state(obj: object) -> bool:
    return obj.__class__.__name__ == "CustomState"


def _is_fock(obj: object) -> bool:
    return obj.__class__.__name__ == "Fock"


def _is_polarization(obj: object) -> bool:
    return obj.__class__.__name__ == "Polarization"


class ProductState:
    """One tensor-product space: the array in HBM plus the subsystems that live in it."""

    __slots__ = ("expansion_level", "container", "uid", "state_objs", "_array", "_layout",
                 "_verified", "_meta_cache", "_meta_key", "_plan_cache")

    def __init__(self, expansion_level: ExpansionLevel, container: "CompositeEnvelopeContainer",
                 state: Optional[Any] = None, state_objs: Optional[List[BaseState]] = None) -> None:
        self.expansion_level = expansion_level
        self.container = container
        self.uid = uuid.uuid4()
        self.state_objs: List[BaseState] = [] if state_objs is None else state_objs
        self._array = dev.scalar_one() if state is None else dev.adopt(state)
        self._layout: List[BaseState] = list(self.state_objs)  # physical axis order of _array
        self._verified: Optional[Any] = None
        self._meta_cache = None
        self._meta_key = None
        self._plan_cache: Dict[Tuple, ShapePlan] = {}

    def __hash__(self) -> int:
        return hash(self.uid)

    # -- the array -----------------------------------------------------------------------
    @property
    def is_matrix(self) -> bool:
        return self.expansion_level == ExpansionLevel.Matrix

    @property
    def state(self) -> Any:
        """The product state with its subsystems in ``state_objs`` order (materialised on
        demand when the physical order differs)."""
        self._sync_layout()
        if len(self._layout) and [id(s) for s in self._layout] != [id(s) for s in self.state_objs]:
            perm = [self._position(s) for s in self.state_objs]
            dims = [s.dimensions for s in self._layout]
            self._array = dev.permute(self._array, dims, perm, self.is_matrix)
            self._layout = list(self.state_objs)
        return self._array

    @state.setter
    def state(self, value: Any) -> None:
        self._array = dev.adopt(value)
        self._layout = list(self.state_objs)

    def _sync_layout(self) -> None:
        """Subsystems that left ``state_objs`` through a measurement have also left the array."""
        if len(self._layout) != len(self.state_objs):
            alive = {id(s) for s in self.state_objs}
            self._layout = [s for s in self._layout if id(s) in alive]

    def _position(self, state: BaseState) -> int:
        for i, s in enumerate(self._layout):
            if s is state:
                return i
        raise ValueError("state is not part of this product state")

    def _physical(self, *targets: BaseState) -> Tuple[List[int], List[int]]:
        """(dims of the array's axes, positions of the targets) in the physical order."""
        self._sync_layout()
        return [s.dimensions for s in self._layout], [self._position(t) for t in targets]

    def _set_array(self, value: Any) -> None:
        """Rebind the array without touching the physical order."""
        self._array = value

    @property
    def size(self) -> int:
        return dev.nbytes(self._array)

    @property
    def is_empty(self) -> bool:
        return len(self.state_objs) == 0

    # -- representation ------------------------------------------------------------------
    def expand(self) -> None:
        """Vector -> Matrix (composite_envelope.py:109-126)."""
        if self.expansion_level < ExpansionLevel.Matrix:
            dims = int(np.prod([s.dimensions for s in self.state_objs]))
            expand = state_expand_jit if Config().use_jit else state_expand
            self._array, self.expansion_level = expand(self._array, self.expansion_level, dims)
            self._array = dev.adopt(self._array)
            for state in self.state_objs:
                state.expansion_level = ExpansionLevel.Matrix

    def contract(self, tol: float = 1e-6) -> None:
        """Matrix -> Vector when the state is pure (composite_envelope.py:128-145)."""
        if self.expansion_level is ExpansionLevel.Matrix:
            self._array, self.expansion_level, success = state_contract(self._array,
                                                                        self.expansion_level)
            if success:
                for state in self.state_objs:
                    state.expansion_level = ExpansionLevel.Vector

    def prepare_meta(self, *target_states: BaseState):
        targets = tuple(target_states) if target_states else tuple(self.state_objs)
        key = (tuple(s.dimensions for s in self.state_objs), tuple(id(t) for t in targets))
        if self._meta_cache is None or self._meta_key != key:
            self._meta_cache = shape_planning.build_meta(self.state_objs, targets)
            self._meta_key = key
        return self._meta_cache

    def prepare_plan(self, *target_states: BaseState) -> ShapePlan:
        """ShapePlan over the LOGICAL order (what ``.state`` exposes)."""
        targets = tuple(target_states) if target_states else tuple(self.state_objs)
        key = (tuple(s.dimensions for s in self.state_objs), tuple(id(s) for s in self.state_objs),
               tuple(id(t) for t in targets))
        if key not in self._plan_cache:
            self._plan_cache[key] = shape_planning.build_plan(self.state_objs, targets)
        return self._plan_cache[key]

    def compiled_kernels(self, *target_states: BaseState):
        return compiled_kernels(self.prepare_plan(*target_states))

    def reorder(self, *ordered_states: BaseState) -> None:
        """New tensor order; all subsystems must be given (composite_envelope.py:201-248).
        Only the logical order changes -- see the module docstring."""
        assert all(so in ordered_states for so in self.state_objs), "All state objects need to be given"
        self._sync_layout()
        self.state_objs = list(ordered_states)
        self.container.update_all_indices()

    # -- measurement ---------------------------------------------------------------------
    def measure(self, *states: BaseState, separate_measurement: bool = False,
                destructive: bool = True, key: Any = None, return_key: bool = False):
        """Joint projective measurement of ``states`` (composite_envelope.py:250-343)."""
        if key is None:
            raise ValueError("PRNG key is required when measuring a composite product state")
        assert all(so in self.state_objs for so in states), "All state objects need to be in product state"
        dims, tgt = self._physical(*states)
        use_key, _ = borrow_key(key)
        fn = measure_matrix_jit if self.is_matrix else measure_vector_jit
        outcomes, post, next_key = fn(self._layout, states, self._array, use_key,
                                      meta=shape_planning.make_meta(tuple(dims), tuple(tgt)))
        self._array = post
        for state in states:
            if destructive:
                state._set_measured()
            else:
                if _is_polarization(state):
                    from .polarization import PolarizationLabel

                    state.state = PolarizationLabel.H if outcomes[state] == 0 else PolarizationLabel.V
                else:
                    state.state = outcomes[state]
                state.expansion_level = ExpansionLevel.Label
                state.index = None
        self.state_objs = [s for s in self.state_objs if s not in states]
        self._sync_layout()
        if Config().contractions and len(self.state_objs) > 0:
            self.contract()
        if return_key:
            return outcomes, next_key
        return outcomes

    def measure_expectation(self, *states: BaseState) -> Tuple[Any, Any]:
        """(probabilities, expected post-measurement density matrix of the rest)
        (composite_envelope.py:345-367)."""
        assert all(so in self.state_objs for so in states), "All state objects need to be in product state"
        # the expected post-measurement matrix keeps the remaining subsystems in tensor order:
        # materialise the logical order so that this is the order the caller sees
        array = self.state
        plan = self.prepare_plan(*states)
        fn = measure_matrix_expectation if self.is_matrix else measure_vector_expectation
        return fn(self.state_objs, states, array, meta=plan)

    def measure_POVM(self, operators: List[Any], *states: BaseState, destructive: bool = True,
                     key: Any = None) -> Tuple[int, Dict[BaseState, int]]:
        """composite_envelope.py:369-426."""
        if self.expansion_level == ExpansionLevel.Vector:
            self.expand()
        dims, tgt = self._physical(*states)
        outcome, post = measure_POVM_matrix(self._layout, states, operators, self._array, key,
                                            meta=shape_planning.make_meta(tuple(dims), tuple(tgt)))
        self._array = post
        outcome = int(outcome)
        other_outcomes: Dict[BaseState, int] = {}
        if destructive:
            for s in states:  # a CustomState cannot be destroyed: it keeps its reduced state
                if _is_custom_state(s):
                    d, t = self._physical(s)
                    s.state = adapters.trace_out_matrix(d, t, self._array)
                    s.expansion_level = ExpansionLevel.Matrix
                    s.index = None
            remaining = [s for s in self.state_objs if s not in states]
            if isinstance(CompositeEnvelope._instances[self.container.composite_uid], list):
                other_outcomes = CompositeEnvelope._instances[self.container.composite_uid][0].measure(
                    *states, key=key)
                for s in states:
                    del other_outcomes[s]
            self.state_objs = remaining
            self._sync_layout()
        if Config().contractions and len(self.state_objs) > 0:
            self.contract()
        return (outcome, other_outcomes)

    # -- channels, partial trace ---------------------------------------------------------
    def apply_kraus(self, operators: Union[List[Any], Tuple[Any, ...]], *states: BaseState) -> None:
        """composite_envelope.py:428-456."""
        while self.expansion_level < ExpansionLevel.Matrix:
            self.expand()
        dims, tgt = self._physical(*states)
        self._array = apply_kraus_matrix(self._layout, states, self._array, operators,
                                         use_contraction=Config().contractions,
                                         meta=shape_planning.make_meta(tuple(dims), tuple(tgt)))
        self._verified = None
        if Config().contractions:
            self.contract()

    def trace_out(self, *states: BaseState) -> Any:
        """Reduced density matrix of ``states``, tensored in the order given
        (composite_envelope.py:458-480)."""
        dims, tgt = self._physical(*states)
        if self.is_matrix:
            return adapters.trace_out_matrix(dims, tgt, self._array)
        return adapters.trace_out_vector(dims, tgt, self._array)

    def occupancy(self) -> Tuple[Dict[int, int], float, float]:
        """ONE pass over the array: ({id(subsystem): highest occupied level}, norm^2 or trace,
        max |entry|^2)."""
        dims, _ = self._physical()
        levels, total, biggest = dev.occupancy(self._array, dims, self.is_matrix)
        return {id(s): lv for s, lv in zip(self._layout, levels)}, total, biggest

    def resize_fock(self, new_dimensions: int, fock: BaseState,
                    occupied: Optional[int] = None) -> bool:
        """Pad the Fock axis, or truncate it when no weight is lost
        (composite_envelope.py:495-585)."""
        if self.expansion_level not in (ExpansionLevel.Vector, ExpansionLevel.Matrix):
            return False
        if new_dimensions == fock.dimensions:
            return False
        dims, (axis,) = self._physical(fock)
        if new_dimensions < fock.dimensions:
            if occupied is None:
                occupied = self.occupancy()[0][id(fock)]
            if occupied >= new_dimensions:
                return False
        self._array = dev.resize(self._array, dims, axis, new_dimensions, self.is_matrix)
        self._verified = None
        fock.dimensions = new_dimensions
        return True

    # -- operations ----------------------------------------------------------------------
    def apply_operation(self, operation: Operation, *states: BaseState) -> None:
        """composite_envelope.py:587-700."""
        if self._verified is not self._array:
            _, biggest = dev.norm_stats(self._array)
            if not biggest > 0:
                raise ValueError("The state is invalidState 0")
        C = Config()
        if C.use_jit and C.dynamic_dimensions:
            raise ValueError("JIT mode does not support dynamic dimension resizing; "
                             "pre-size states or disable use_jit.")
        kind = operation._operation_type
        if isinstance(kind, FockOperationType):
            if not _is_fock(states[0]):
                raise ValueError("Fock operation requires a Fock state")
            assert len(states) == 1
            if C.dynamic_dimensions:
                operation.compute_dimensions(states[0]._num_quanta, states[0].trace_out())
                states[0].resize(operation.dimensions[0])
            else:
                operation.dimensions = [states[0].dimensions]
        elif isinstance(kind, PolarizationOperationType):
            if not _is_polarization(states[0]):
                raise ValueError("Polarization operation requires a Polarization state")
            assert len(states) == 1
            operation.compute_dimensions(0, None)
        elif isinstance(kind, CustomStateOperationType):
            if not _is_custom_state(states[0]):
                raise ValueError("CustomState operation requires a CustomState")
            assert len(states) == 1
            operation.compute_dimensions(0, states[0].trace_out())
        elif isinstance(kind, CompositeOperationType):
            assert len(states) == len(kind.expected_base_state_types)
            for s, expected in zip(states, kind.expected_base_state_types):
                assert isinstance(s, expected)
            if C.dynamic_dimensions:
                operation.compute_dimensions(
                    [s._num_quanta if _is_fock(s) else s.dimensions for s in states],
                    [s.trace_out() for s in states])
                for i, s in enumerate(states):
                    if _is_fock(s):
                        s.resize(operation._dimensions[i])
            else:
                operation.dimensions = [s.dimensions for s in states]

        dims, tgt = self._physical(*states)
        apply = apply_operation_matrix if self.is_matrix else apply_operation_vector
        self._array = apply(self._layout, states, self._array, operation.device_operator,
                            meta=shape_planning.make_meta(tuple(dims), tuple(tgt)),
                            use_contraction=C.contractions)

        # one pass: all-zero test, norm, and the highest occupied level of every subsystem
        levels, total, biggest = self.occupancy()
        if not biggest > 0:
            raise ValueError("The state is entirely composed of zeros, "
                             "is |0⟩ attempted to be annihilated?")
        if operation.renormalize:
            if self.is_matrix:
                # the reference divides rho by its Frobenius norm (composite_envelope.py:684-685)
                total, _ = dev.norm_stats(self._array)
            self._array = dev.divide(self._array, math.sqrt(total))
        # shrink every Fock space to its support (composite_envelope.py:687-696); the
        # reference's resize_fock moves each Fock space to the front on the way
        for so in list(self.state_objs):
            if _is_fock(so):
                top = max(1, levels[id(so)])
                composite = CompositeEnvelope._instances[self.container.composite_uid][0]
                composite.reorder(so)
                self.resize_fock(top + 1, so, occupied=levels[id(so)])
        self._verified = self._array
        if Config().contractions:
            self.contract()


@dataclass
class CompositeEnvelopeContainer:
    composite_uid: uuid.UUID
    envelopes: List[Envelope] = field(default_factory=list)
    state_objs: List[BaseState] = field(default_factory=list)
    states: List[ProductState] = field(default_factory=list)

    def append_states(self, other: "CompositeEnvelopeContainer") -> None:
        assert isinstance(other, CompositeEnvelopeContainer)
        self.states.extend(other.states)
        self.envelopes.extend(other.envelopes)

    def remove_empty_product_states(self) -> None:
        for state in self.states[:]:
            if state is not None and state.is_empty:
                self.states.remove(state)

    def update_all_indices(self) -> None:
        """Every subsystem points at (product state, logical position)."""
        for state_index, state in enumerate(self.states):
            for i, so in enumerate(state.state_objs):
                if so is not None:
                    so.extract((state_index, i))
                    so.composite_envelope = CompositeEnvelope._instances[self.composite_uid][0]


class CompositeEnvelope:
    """A pointer to a container; several CompositeEnvelopes can share one container
    (composite_envelope.py:757-1457)."""

    _containers: Dict[Union[str, uuid.UUID], CompositeEnvelopeContainer] = {}
    _instances: Dict[Union[str, uuid.UUID], List["CompositeEnvelope"]] = {}

    def __init__(self, *states: Any):
        self.uid = uuid.uuid4()
        composites: List[CompositeEnvelope] = [e for e in states if isinstance(e, CompositeEnvelope)]
        envelopes = [e for e in states if _is_envelope(e)]
        state_objs: List[BaseState] = []
        for e in states:
            if _is_custom_state(e):
                state_objs.append(e)
            elif _is_envelope(e):
                state_objs.extend([e.fock, e.polarization])
        for e in envelopes:
            if e.composite_envelope is not None and e.composite_envelope not in composites:
                composites.append(e.composite_envelope)
        container: Optional[CompositeEnvelopeContainer] = None
        for ce in composites:
            assert isinstance(ce, CompositeEnvelope), "ce should be CompositeEnvelope type"
            state_objs.extend(ce.state_objs)
            if container is None:
                container = CompositeEnvelope._containers[ce.uid]
            else:
                container.append_states(CompositeEnvelope._containers[ce.uid])
            ce.uid = self.uid
        if container is None:
            container = CompositeEnvelopeContainer(self.uid)
        for e in envelopes:
            if e not in container.envelopes:
                container.envelopes.append(e)
        known = {x.uid for x in container.state_objs}
        for s in state_objs:
            if s.uid not in known:
                container.state_objs.append(s)
                known.add(s.uid)
        CompositeEnvelope._containers[self.uid] = container
        CompositeEnvelope._instances.setdefault(self.uid, []).append(self)
        self.update_composite_envelope_pointers()

    def __repr__(self) -> str:
        return (f"CompositeEnvelope(uid={self.uid}, envelopes={[e.uid for e in self.envelopes]},"
                f"state_objects={[s.uid for s in self.state_objs]})")

    @property
    def envelopes(self) -> List[Envelope]:
        return CompositeEnvelope._containers[self.uid].envelopes

    @property
    def state_objs(self) -> List[BaseState]:
        return CompositeEnvelope._containers[self.uid].state_objs

    @property
    def product_states(self) -> List[ProductState]:
        return CompositeEnvelope._containers[self.uid].states

    @property
    def container(self) -> CompositeEnvelopeContainer:
        return CompositeEnvelope._containers[self.uid]

    @property
    def states(self) -> List[ProductState]:
        return CompositeEnvelope._containers[self.uid].states

    def update_composite_envelope_pointers(self) -> None:
        for envelope in self.envelopes:
            envelope.set_composite_envelope_id(self.uid)

    def _holding(self, states: Tuple[BaseState, ...]) -> List[ProductState]:
        return [p for p in self.states if any(so in p.state_objs for so in states)]

    def _single_space(self, states: Tuple[BaseState, ...]) -> ProductState:
        """The product state holding ``states``, merging spaces when they are spread out."""
        holding = self._holding(states)
        if len(holding) > 1:
            everything = list(states)
            for p in holding:
                everything.extend(p.state_objs)
            self.combine(*everything)
        elif len(holding) == 0:
            self.combine(*states)
        return self._holding(states)[0]

    def expand(self, *states: BaseState) -> None:
        for p in self._holding(states):
            p.expand()

    def contract(self, *state_objs: BaseState) -> None:
        pass

    def combine(self, *state_objs: BaseState) -> None:
        """Bring the given subsystems into ONE product state: Kronecker product of the
        product states / envelope arrays / free states that hold them
        (composite_envelope.py:869-963)."""
        wanted = set(state_objs)
        for ps in self.states:
            if wanted.issubset(ps.state_objs):
                return
        assert all(s in self.state_objs for s in state_objs)
        existing: List[ProductState] = []
        for state in state_objs:
            for ps in self.product_states:
                if state in ps.state_objs:
                    existing.append(ps)
        existing = list(dict.fromkeys(existing))

        level = ExpansionLevel.Vector
        for obj in state_objs:
            assert isinstance(obj.expansion_level, ExpansionLevel)
            if obj.expansion_level > level:
                level = ExpansionLevel.Matrix
                break
        for ps in existing:
            while ps.expansion_level < level:
                ps.expand()
        for obj in state_objs:
            if obj.index is None or (isinstance(obj.index, int) and hasattr(obj, "envelope")):
                while obj.expansion_level < level:
                    obj.expand()

        arrays: List[Any] = []
        order: List[BaseState] = []
        for ps in existing:
            arrays.append(ps.state)  # logical order
            order.extend(ps.state_objs)
            ps.state_objs = []
            ps.state = dev.scalar_one()
        for so in state_objs:
            env = getattr(so, "envelope", None)
            if (env is not None and so.index is not None and env.state is not None
                    and not isinstance(so.index, tuple)):
                arrays.append(env.state)
                slots: List[Optional[BaseState]] = [None, None]
                if isinstance(env.fock.index, int):
                    slots[env.fock.index] = env.fock
                if isinstance(env.polarization.index, int):
                    slots[env.polarization.index] = env.polarization
                order.extend([s for s in slots if s is not None])
                env.state = None
            elif so.index is None:
                assert dev.is_array(so.state)
                arrays.append(so.state)
                order.append(so)
                so.state = None
        ps = ProductState(expansion_level=level, container=self._containers[self.uid],
                          state=kron_reduce(arrays), state_objs=order)
        CompositeEnvelope._containers[self.uid].states.append(ps)
        self.container.remove_empty_product_states()
        self.container.update_all_indices()

    def reorder(self, *ordered_states: BaseState) -> None:
        """Put the given subsystems, in this order, at the front of their product state by
        pairwise swaps (composite_envelope.py:965-1007)."""
        self.combine(*ordered_states)
        ps = [p for p in self.states if all(so in p.state_objs for so in ordered_states)][0]
        new_order = list(ps.state_objs)
        for i, wanted in enumerate(ordered_states):
            j = new_order.index(wanted)
            if j != i:
                new_order[i], new_order[j] = new_order[j], new_order[i]
        ps.reorder(*new_order)

    def measure(self, *states: BaseState, separate_measurement: bool = True, destructive: bool = True,
                key: Any = None, return_key: bool = False):
        """composite_envelope.py:1009-1139."""
        C = Config()
        if key is None and C.use_jit:
            raise ValueError("PRNG key is required when Config.use_jit is True for composite"
                             "measurements")
        if key is None:
            key = C.random_key
        outcomes: Dict[BaseState, int] = {}
        current_key = key
        next_key = None
        state_list = list(states)
        if not separate_measurement:
            for s in state_list:
                env = getattr(s, "envelope", None)
                if env is not None and _is_envelope(env):
                    partner = env.polarization if _is_fock(s) else env.fock if _is_polarization(s) else None
                    if partner is not None and partner not in state_list:
                        state_list.append(partner)
        # subsystems that still live in their own object or envelope are measured there
        for s in state_list:
            if isinstance(s.index, int) and hasattr(s, "envelope"):
                env = s.envelope
                if env is None or not _is_envelope(env):
                    raise ValueError("State is missing envelope for measurement")
                use_key, current_key = borrow_key(current_key)
                if separate_measurement:
                    out, current_key = env.measure(s, separate_measurement=True, destructive=destructive,
                                                   key=use_key, return_key=True)
                else:
                    partner = env.fock if _is_polarization(s) else env.polarization
                    out, current_key = env.measure(s, partner, separate_measurement=False,
                                                   destructive=destructive, key=use_key, return_key=True)
                outcomes.update(out)
            elif s.index is None and not s.measured:
                use_key, current_key = borrow_key(current_key)
                outcomes.update(s.measure(separate_measurement=separate_measurement,
                                          destructive=destructive, key=use_key))
        for ps in self._holding(tuple(state_list)):
            ps_states = [so for so in state_list if so in ps.state_objs]
            use_key, current_key = borrow_key(current_key)
            out, next_key = ps.measure(*ps_states, separate_measurement=separate_measurement,
                                       destructive=destructive, key=use_key, return_key=True)
            outcomes.update(out)
        if destructive:
            for s in state_list:
                env = getattr(s, "envelope", None)
                if env is not None and _is_envelope(env):
                    env._set_measured()
        self._containers[self.uid].update_all_indices()
        self._containers[self.uid].remove_empty_product_states()
        if return_key:
            return outcomes, next_key
        return outcomes

    def measure_expectation(self, *states: BaseState) -> Tuple[Any, Any]:
        """composite_envelope.py:1141-1177."""
        if len(states) == 0:
            assert len(self.product_states) > 0
            ps = self.product_states[0]
            return ps.measure_expectation(*ps.state_objs)
        return self._single_space(states).measure_expectation(*states)

    def _check_operator_shapes(self, operators: List[Any], states: Tuple[BaseState, ...]) -> None:
        dim = int(np.prod([s.dimensions for s in states]))
        for op in operators:
            if tuple(op.shape) != (dim, dim):
                raise ValueError("At least on Kraus operator has incorrect dimensions: "
                                 f"{tuple(op.shape)}, expected({dim},{dim})")

    def measure_POVM(self, operators: List[Any], *states: BaseState, destructive: bool = True,
                     key: Any = None) -> Tuple[int, Dict[BaseState, int]]:
        """composite_envelope.py:1179-1238."""
        self._check_operator_shapes(operators, states)
        ps = self._single_space(states)
        self.reorder(*states)
        return ps.measure_POVM(operators, *states, destructive=destructive, key=key)

    def apply_kraus(self, operators: List[Any], *states: BaseState, identity_check: bool = True) -> None:
        """composite_envelope.py:1240-1315."""
        if len(states) != len(set(states)):
            raise ValueError("State list should contain unique elements")
        self._check_operator_shapes(operators, states)
        if identity_check and not kraus_identity_check(operators):
            raise ValueError("Kraus operators do not sum to the identity")
        if len(self._holding(states)) == 0:
            # not combined yet: apply where the state lives
            if len(states) == 1:
                states[0].apply_kraus(operators)
                return
            if len(states) == 2 and hasattr(states[0], "envelope") and hasattr(states[1], "envelope"):
                env = states[0].envelope
                if env is states[1].envelope and env is not None and _is_envelope(env):
                    env.apply_kraus(operators, *states)
                    return
        ps = self._single_space(states)
        self.reorder(*states)
        ps.apply_kraus(operators, *states)

    def trace_out(self, *states: BaseState) -> Any:
        """composite_envelope.py:1317-1355."""
        holding = self._holding(states)
        assert len(holding) > 0, "No product state found"
        ps = self._single_space(states)
        if len(states) > 1:
            self.reorder(*states)
        return ps.trace_out(*states)

    def resize_fock(self, new_dimensions: int, fock: BaseState) -> bool:
        """composite_envelope.py:1357-1394."""
        if not _is_fock(fock):
            raise ValueError("Only Fock spaces can be resized")
        if fock not in self.state_objs:
            raise ValueError("Tried to resizing fock, which is not a part of this envelope")
        if not isinstance(fock.index, tuple):
            return fock.resize(new_dimensions)
        ps = [ps for ps in self.product_states if fock in ps.state_objs]
        if len(ps) != 1:
            raise ValueError("Something went wrong")
        self.reorder(fock)
        return ps[0].resize_fock(new_dimensions, fock)

    def apply_operation(self, operator: Operation, *states: BaseState) -> None:
        """Route to the object / envelope while the state lives there, otherwise to the (merged)
        product state (composite_envelope.py:1396-1434)."""
        if len(states) == 1 and not isinstance(states[0].index, tuple):
            states[0].apply_operation(operator)
            return
        holding = self._holding(states)
        if len(holding) == 1 and all(s in holding[0].state_objs for s in states):
            ps = holding[0]
        else:
            everything = list(states)
            for p in holding:
                everything.extend(p.state_objs)
            self.combine(*everything)
            ps = self._holding(states)[0]
        ps.apply_operation(operator, *states)

    def prepare_meta(self, product_state_index: int = 0, *target_states: BaseState):
        ps = self.product_states[product_state_index]
        return ps.prepare_meta(*target_states)

    def compiled_kernels(self, product_state_index: int = 0, *target_states: BaseState):
        ps = self.product_states[product_state_index]
        return ps.compiled_kernels(*target_states)
