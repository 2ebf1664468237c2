"""
``Operation``: an operation type plus its parameters
(reference: photon_weave/operation/operation.py:17-335).

Operator matrices are built on the host (they are t x t) and cached, host copy and HBM
copy, under ``(type, dimensions, parameters)``: the reference recomputes ``expm`` / ``kron``
on every ``.operator`` access (operation/operation.py:292-294), which for a Jaynes-Cummings
style loop is most of the step time.
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np

from ..core import device as dev
from ..state.expansion_levels import ExpansionLevel
from .composite_operation import CompositeOperationType
from .custom_state_operation import CustomStateOperationType
from .fock_operation import FockOperationType
from .polarization_operation import PolarizationOperationType

OperationType = Union[FockOperationType, PolarizationOperationType, CompositeOperationType,
                      CustomStateOperationType]

_CACHE: Dict[Tuple, Tuple[np.ndarray, Any]] = {}
_CACHE_LIMIT = 512


def _freeze(value: Any) -> Any:
    """Hashable fingerprint of an operator parameter; None when it cannot be cached safely."""
    if isinstance(value, (int, float, complex, str, bool, type(None))):
        return value
    if isinstance(value, np.generic):
        return value.item()
    if isinstance(value, (tuple, list)):
        parts = tuple(_freeze(v) for v in value)
        return None if any(p is None and v is not None for p, v in zip(parts, value)) else parts
    return None  # arrays, callables, dict contexts: evaluated every time


class Operation:
    __slots__ = ("_operator", "_operation_type", "_apply_count", "_renormalize", "kwargs",
                 "_expansion_level", "_expression", "_dimensions", "_device_operator")

    def __init__(self, operation_type: OperationType, expression: Optional[str] = None,
                 apply_count: int = 1, **kwargs: Any) -> None:
        self._operation_type = operation_type
        self._operator: Optional[Any] = None
        self._device_operator: Optional[Any] = None
        self._apply_count = apply_count
        self.kwargs = kwargs
        self._operation_type.update(**kwargs)
        if operation_type is FockOperationType.Custom:
            op = kwargs["operator"]
            assert hasattr(op, "shape") and len(op.shape) == 2
            self._operator = op
            self._dimensions: List[int] = [int(op.shape[0])]
        for param in operation_type.required_params:
            if param not in kwargs:
                raise KeyError(f"The '{param}' argument is required for {operation_type.name}")

    def __repr__(self) -> str:
        """operation.py:138-176."""
        head = f"{self._operation_type.__class__.__name__}.{self._operation_type.name}"
        if self._operator is None:
            return head
        rows = []
        for row in dev.to_numpy(self._operator):
            cells = "".join(
                f"{complex(v).real:+.2f} {'+' if complex(v).imag >= 0 else '-'} {abs(complex(v).imag):.2f}j   "
                for v in row)
            rows.append(("⎢ " + cells).strip() + " ⎥")
        rows[0] = "⎡" + rows[0][1:-1] + "⎤"
        rows[-1] = "⎣" + rows[-1][1:-1] + "⎦"
        return head + "\n" + "\n".join(rows)

    @property
    def dimensions(self) -> List[int]:
        assert isinstance(self._dimensions, list)
        return self._dimensions

    @dimensions.setter
    def dimensions(self, dimensions: List[int]) -> None:
        self._dimensions = dimensions

    def compute_dimensions(self, num_quanta: Union[int, List[int]], state: Any) -> None:
        """operation.py:205-224."""
        if self._operation_type is not FockOperationType.Custom:
            self._dimensions = self._operation_type.compute_dimensions(num_quanta, state, **self.kwargs)

    @property
    def required_expansion_level(self) -> ExpansionLevel:
        return self._operation_type.required_expansion_level

    @property
    def renormalize(self) -> bool:
        return self._operation_type.renormalize

    def _cache_key(self) -> Optional[Tuple]:
        frozen = []
        for name in sorted(self.kwargs):
            f = _freeze(self.kwargs[name])
            if f is None and self.kwargs[name] is not None:
                return None
            frozen.append((name, f))
        return (self._operation_type, tuple(int(d) for d in self.dimensions), tuple(frozen))

    def _materialise(self) -> Tuple[Any, Any]:
        key = self._cache_key()
        if key is not None and key in _CACHE:
            return _CACHE[key]
        host = self._operation_type.compute_operator(self.dimensions, **self.kwargs)
        if dev.is_array(host):  # user-supplied device operator
            return (host, host)
        host = np.asarray(host)
        if key is None and host.nbytes <= (1 << 16):
            # parameters that cannot be fingerprinted (arrays, callables): key on the matrix itself
            key = ("bytes", host.shape, str(host.dtype), host.tobytes())
            if key in _CACHE:
                return _CACHE[key]
        pair = (host, dev.asarray(host))
        if key is not None:
            # the cached device operator is immutable for as long as it lives: let libpwb200 keep
            # its structure analysis (block structure, super-blocks, superoperator) across applies
            dev.pin_operator(pair[1])
            if len(_CACHE) >= _CACHE_LIMIT:
                _CACHE.pop(next(iter(_CACHE)))
            _CACHE[key] = pair
        return pair

    @property
    def operator(self) -> Any:
        """The operator matrix for the current ``dimensions`` (host array)."""
        self._operator, self._device_operator = self._materialise()
        assert hasattr(self._operator, "shape")
        return self._operator

    @operator.setter
    def operator(self, operator: Any) -> None:
        assert hasattr(operator, "shape")
        if self._operation_type is not FockOperationType.Custom:
            raise ValueError("Operator can only be configured for the Custom types")
        self._operator = operator
        self.kwargs["operator"] = operator

    @property
    def device_operator(self) -> Any:
        """Same matrix, resident in HBM (what the kernels read)."""
        self._operator, self._device_operator = self._materialise()
        return self._device_operator

    def to_spec(self, targets: Tuple[int, ...], rep: str, use_contraction: Optional[bool] = None):
        """Pure-data description of this operation (operation.py:318-335)."""
        return {"name": str(self._operation_type), "operator": self.operator, "targets": tuple(targets),
                "rep": rep, "use_contraction": use_contraction}
