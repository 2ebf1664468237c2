"""
Common behaviour of Fock / Polarization / CustomState
(reference: photon_weave/state/base_state.py:36-323).  ``state`` is an ``int`` /
``PolarizationLabel`` at Label level and a CUDA array -- (d, 1) or (d, d) -- above it.
"""

from __future__ import annotations

import math
import sys
from abc import ABC, abstractmethod
from enum import Enum
from typing import Any, Dict, List, Optional, Tuple, Union
from uuid import UUID, uuid4

from ..core import device as dev
from ..core.ops import kraus_identity_check
from ..photon_weave import Config
from .expansion_levels import ExpansionLevel
from .utils.measurements import measure_POVM_matrix
from .utils.operations import apply_kraus_matrix
from .utils.representation import representation_matrix, representation_vector
from .utils.routing import route_operation


def _is_envelope(obj: object) -> bool:
    return hasattr(obj, "fock") and hasattr(obj, "polarization")


def _is_polarization(obj: object) -> bool:
    return obj.__class__.__name__ == "Polarization"


def check_and_renormalize(state: Any, renormalize: bool, message: str) -> Any:
    """Post-apply bookkeeping of the single-state containers (fock.py:486-495,
    polarization.py:386-393, custom_state.py:391-398): one reduction gives both the
    all-zero test and the norm; the reference's ``state / jnp.linalg.norm(state)`` is the
    2-norm for a vector and the Frobenius norm for a matrix."""
    total, biggest = dev.norm_stats(state)
    if not biggest > 0:
        raise ValueError(message)
    if renormalize:
        state = dev.divide(state, math.sqrt(total))
    return state


class BaseState(ABC):
    __slots__ = ("_uid", "_expansion_level", "_index", "_dimensions", "_measured",
                 "_composite_envelope", "_envelope", "_state")

    @abstractmethod
    def __init__(self) -> None:
        self._uid: Union[str, UUID] = uuid4()
        self._expansion_level: Optional[ExpansionLevel] = ExpansionLevel.Label
        self._index: Optional[Union[int, Tuple[int, int]]] = None
        self._dimensions: int = -1
        self._composite_envelope: Optional[Any] = None
        self._envelope: Optional[Any] = None
        self._state: Optional[Union[int, Any]] = None

    # -- plain attributes -------------------------------------------------------------
    @property
    def state(self) -> Optional[Union[int, Any]]:
        return self._state

    @state.setter
    def state(self, value: Optional[Union[int, Any]]) -> None:
        self._state = dev.adopt(value)

    @property
    def size(self) -> int:
        """Bytes held by this object (base_state.py:59-74)."""
        if self.state is None:
            return 0
        if dev.is_array(self.state):
            return dev.nbytes(self.state)
        return sys.getsizeof(self.state) if self.state else 0

    @property
    def envelope(self) -> Any:
        return self._envelope

    @envelope.setter
    def envelope(self, envelope: Any) -> None:
        self._envelope = envelope

    @property
    def measured(self) -> bool:
        return self._measured

    @measured.setter
    def measured(self, measured: bool) -> None:
        self._measured = measured

    @property
    def composite_envelope(self) -> Any:
        return self._composite_envelope

    @composite_envelope.setter
    def composite_envelope(self, composite_envelope: Any) -> None:
        self._composite_envelope = composite_envelope

    @property
    def uid(self) -> Union[str, UUID]:
        return self._uid

    @uid.setter
    def uid(self, uid: Union[UUID, str]) -> None:
        self._uid = uid

    @property
    def dimensions(self) -> int:
        return self._dimensions

    @dimensions.setter
    def dimensions(self, dimensions: int) -> None:
        self._dimensions = dimensions

    @property
    def expansion_level(self) -> Optional[ExpansionLevel]:
        return self._expansion_level

    @expansion_level.setter
    def expansion_level(self, expansion_level: Optional[ExpansionLevel]) -> None:
        self._expansion_level = expansion_level

    @property
    def index(self) -> Union[None, int, Tuple[int, int]]:
        return self._index

    @index.setter
    def index(self, index: Union[None, int, Tuple[int, int]]) -> None:
        self._index = index

    def __hash__(self) -> int:
        return hash(self.uid)

    def __repr__(self) -> str:
        """base_state.py:136-153."""
        if self.expansion_level == ExpansionLevel.Label:
            if isinstance(self.state, Enum):
                return f"|{self.state.value}⟩"
            if isinstance(self.state, int):
                return f"|{self.state}⟩"
        if self.index is not None or self.measured:
            return str(self.uid)
        if self.expansion_level == ExpansionLevel.Vector:
            return representation_vector(self.state)
        if self.expansion_level == ExpansionLevel.Matrix:
            return representation_matrix(self.state)
        return f"{self.uid}"

    # -- shared operations ------------------------------------------------------------
    def apply_kraus(self, operators: List[Any], identity_check: bool = True) -> None:
        """base_state.py:155-194: the state is expanded to a density matrix first."""
        if identity_check and not kraus_identity_check(operators):
            raise ValueError("Invalid Kraus Channel")
        while self.expansion_level < ExpansionLevel.Matrix:
            self.expand()
        self.state = apply_kraus_matrix([self], [self], self.state, operators,
                                        use_contraction=Config().contractions)
        if Config().contractions:
            self.contract()

    @abstractmethod
    def expand(self) -> None: ...

    @abstractmethod
    def _set_measured(self) -> None: ...

    @abstractmethod
    def extract(self, index: Union[int, Tuple[int, int]]) -> None: ...

    def contract(self, final: ExpansionLevel = ExpansionLevel.Label, tol: float = 1e-6) -> None:
        raise NotImplementedError("contract must be implemented by concrete state classes")

    @route_operation()
    def measure_POVM(self, operators: List[Any], destructive: bool = True, partial: bool = False,
                     key: Any = None) -> Tuple[int, Dict["BaseState", int]]:
        """base_state.py:226-291."""
        while self.expansion_level < ExpansionLevel.Matrix:
            self.expand()
        assert dev.shape(self.state) == (self.dimensions, self.dimensions)
        outcome, self.state = measure_POVM_matrix([self], [self], operators, self.state, key)
        outcome = int(outcome)
        result: Tuple[int, Dict["BaseState", int]] = (outcome, {})
        if destructive:
            self._set_measured()
        if not partial:
            env = self.envelope
            if env is not None and _is_envelope(env):
                other = env.fock if _is_polarization(self) else env.polarization
                for k, v in other.measure(destructive=destructive, key=key).items():
                    result[1][k] = v
        if Config().contractions and not destructive:
            self.contract()
        return result

    @route_operation()
    def trace_out(self) -> Any:
        """The reduced state of this subsystem; routed to the container when the state
        lives in one (base_state.py:293-313)."""
        assert self.state is not None
        return self.state

    @abstractmethod
    def measure(self, separate_measurement: bool = False, destructive: bool = True,
                key: Any = None) -> Dict["BaseState", int]: ...

    @abstractmethod
    def measure_expectation(self) -> Tuple[Any, Any]: ...
