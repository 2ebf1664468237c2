"""Pretty printers for small states (reference: photon_weave/state/utils/representation.py)."""

from __future__ import annotations

from typing import Any, List

from ...core import device as dev


def _entry(num: complex, pad: str) -> str:
    sign = "+" if num.imag >= 0 else "-"
    return f"{num.real:+.2f} {sign} {abs(num.imag):.2f}j{pad}"


def _rows(state: Any, pad: str) -> List[str]:
    arr = dev.to_numpy(state)
    return [("⎢ " + "".join(_entry(complex(v), pad) for v in row)).strip() + " ⎥" for row in arr]


def representation_vector(state: Any) -> str:
    """representation.py:6-40."""
    assert dev.shape(state)[1] == 1
    lines = _rows(state, " ")
    lines[0] = "⎡ " + lines[0][2:-1] + "⎤"
    lines[-1] = "⎣ " + lines[-1][2:-1] + "⎦"
    return "\n".join(lines)


def representation_matrix(state: Any) -> str:
    """representation.py:43-86."""
    shp = dev.shape(state)
    assert shp[0] == shp[1]
    lines = _rows(state, "   ")
    lines[0] = "⎡" + lines[0][1:-1] + "⎤"
    lines[-1] = "⎣" + lines[-1][1:-1] + "⎦"
    return "\n".join(lines)
