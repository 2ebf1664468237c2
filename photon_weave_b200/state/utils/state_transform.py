"""
Label <-> Vector <-> Matrix changes of representation
(reference: photon_weave/state/utils/state_transform.py:10-148), on the device:

* Vector -> Matrix is the outer-product kernel ``pw_expand_vec``;
* Matrix -> Vector replaces ``rho @ rho`` + ``eigh`` by ``pw_contract_mat``: one reduction
  for Tr rho^2 and, for a pure state, the normalised column of the largest diagonal entry
  with the global phase fixed so that psi[0] is real and non-negative (the reference's
  convention, state_transform.py:132-135);
* Vector -> Label is ``pw_find_label`` (exactly one entry equal to 1).
"""

from __future__ import annotations

from typing import Any, Tuple, Union

from ...core import device as dev
from ..expansion_levels import ExpansionLevel


def state_expand(state: Union[int, Any], current_expansion_level: ExpansionLevel,
                 dimensions: int) -> Tuple[Any, ExpansionLevel]:
    """state_transform.py:10-66."""
    dimensions = int(dimensions)
    if not isinstance(current_expansion_level, ExpansionLevel):
        current_expansion_level = ExpansionLevel(current_expansion_level)
    if dimensions < 0:
        raise ValueError("Dimensions must be larger than 0")
    if current_expansion_level is ExpansionLevel.Label:
        if not isinstance(state, int):
            raise ValueError("Could not expand state, where label is not int type")
        assert state >= 0
        return dev.basis_vector(dimensions, state), ExpansionLevel.Vector
    state = dev.adopt(state)
    if current_expansion_level is ExpansionLevel.Vector:
        assert dev.is_array(state)
        assert dev.shape(state) == (dimensions, 1)
        return dev.expand(state), ExpansionLevel.Matrix
    assert dev.is_array(state)
    return state, current_expansion_level


def state_expand_jit(state, current_expansion_level, dimensions):
    """Same computation; there is nothing to trace on this path."""
    return state_expand(state, current_expansion_level, dimensions)


def state_contract(state: Union[int, Any], current_expansion_level: ExpansionLevel,
                   tol: float = 1e-6) -> Tuple[Union[int, Any], ExpansionLevel, bool]:
    """state_transform.py:78-148 -> (state, level, success)."""
    assert isinstance(current_expansion_level, ExpansionLevel)
    if current_expansion_level is ExpansionLevel.Matrix:
        assert dev.is_array(state)
        shp = dev.shape(state)
        assert shp[0] == shp[1]
        psi, _, ok = dev.contract(state, tol)
        if ok:
            return psi, ExpansionLevel.Vector, True
    elif current_expansion_level is ExpansionLevel.Vector:
        assert dev.is_array(state)
        assert dev.shape(state)[1] == 1
        count, index = dev.find_label(state)
        if count == 1:
            return int(index), ExpansionLevel.Label, True
    return state, current_expansion_level, False
