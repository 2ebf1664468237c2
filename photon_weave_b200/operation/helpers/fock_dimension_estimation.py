"""
Cut-off estimation for displacement / squeezing / expression operators under
``Config().dynamic_dimensions`` (reference:
photon_weave/operation/helpers/fock_dimension_esitmation.py:9-142).

Works on the REDUCED state of one Fock mode (d or d x d numbers, produced by the trace-out
kernel) and on candidate operators -- operator-sized host arithmetic that decides a shape;
the state itself is never touched here.
"""

from __future__ import annotations

from typing import TYPE_CHECKING, Any

import numpy as np

from ...core import device as dev

if TYPE_CHECKING:
    from ..operation import Operation


class FockDimensions:
    def __init__(self, state: Any, operation: "Operation", num_quanta: int, threshold: float) -> None:
        self.state = np.array(dev.to_numpy(state), dtype=np.complex128)
        self.dimensions = self.state.shape[0]
        self.operation = operation
        self.threshold = threshold
        self.num_quanta = num_quanta

    def compute_dimensions(self) -> int:
        self._initial_estimate()
        while True:
            n = self._compute_dimensions()
            if n > 0:
                return n
            self._increase_dimensions(5)

    def _initial_estimate(self) -> None:
        from ..fock_operation import FockOperationType

        kind = self.operation._operation_type
        if kind is FockOperationType.Displace:
            cutoff = int(self.num_quanta + 3 * np.abs(self.operation.kwargs["alpha"]) ** 2)
            if cutoff > self.dimensions:
                self._increase_dimensions(cutoff - self.dimensions)
        if kind is FockOperationType.Squeeze:
            r = np.abs(self.operation.kwargs["zeta"])
            en = (2 * self.num_quanta + 1) * np.sinh(r) ** 2 + self.num_quanta
            cutoff = int(self.num_quanta + 3 * int(np.ceil(en)))
            if cutoff > self.dimensions:
                self._increase_dimensions(cutoff - self.dimensions)

    def _compute_dimensions(self) -> int:
        d = self.dimensions
        self.operation._dimensions = [d]
        operator = np.asarray(self.operation._operation_type.compute_operator([d], **self.operation.kwargs))
        tail = (1 - self.threshold) * 1e-3
        if self.state.shape == (d, 1):
            out = operator @ self.state
            # the reference compares the complex tail entry itself with the bound
            # (fock_dimension_esitmation.py:71), i.e. its real part
            if out[-1, 0].real > tail:
                return -1
            weights = np.abs(out[:, 0]) ** 2
        elif self.state.shape == (d, d):
            out = operator @ self.state @ operator.T.conj()
            if np.abs(out[-1, -1]) > tail:
                return -1
            weights = np.abs(np.diagonal(out))
        else:
            return -1
        cdf = 0.0
        for i, w in enumerate(weights):
            cdf += float(w)
            if cdf >= self.threshold:
                return i + 3
        return -1

    def _increase_dimensions(self, amount: int = 1) -> None:
        d = self.dimensions
        if self.state.shape == (d, 1):
            self.state = np.vstack([self.state, np.zeros((amount, 1), dtype=self.state.dtype)])
        elif self.state.shape == (d, d):
            self.state = np.pad(self.state, ((0, amount), (0, amount)))
        self.dimensions += amount
