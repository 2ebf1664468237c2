"""
Prefix-tuple expression language for custom operators
(reference: photon_weave/extra/expression_interpreter.py:7-85).  Operator-sized host math.
"""

from __future__ import annotations

from functools import reduce
from typing import Any, Callable, Dict, List

import numpy as np
from scipy.linalg import expm


def interpreter(expr: Any, context: Dict[str, Callable[[List[int]], Any]], dimensions: List[int]) -> Any:
    ev = lambda e: interpreter(e, context, dimensions)  # noqa: E731
    if isinstance(expr, tuple):
        op, *args = expr
        if op == "add":
            return reduce(np.add, [ev(a) for a in args])
        if op == "sub":
            return np.subtract(ev(args[0]), ev(args[1]))
        if op == "s_mult":
            return reduce(lambda x, y: x * y, [ev(a) for a in args])
        if op == "m_mult":
            return reduce(lambda x, y: x @ y, [ev(a) for a in args])
        if op == "kron":
            return reduce(np.kron, [ev(a) for a in args])
        if op == "expm":
            return expm(np.asarray(ev(args[0])))
        if op == "div":
            return ev(args[0]) / ev(args[1])
    elif isinstance(expr, str):
        return context[expr](dimensions)
    else:
        return expr
    raise ValueError("Something went wrong in the expression interpreter!", expr)
