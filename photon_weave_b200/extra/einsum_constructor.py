"""
Einstein-summation strings for product-state manipulations
(reference: photon_weave/extra/einsum_constructor.py:7-370).

Kept for API compatibility: the CUDA kernels address subsystems through strides and never
evaluate these strings, but callers of the reference use them as a *description* of an
operation (and its tests pin their exact text).  One label table drives every generator:
labels are handed out in the order the reference hands them out, ``chr(97 + i)``.
"""

from __future__ import annotations

from typing import Dict, Iterable, List, Sequence


class _Labels:
    """Allocator of consecutive index labels."""

    def __init__(self) -> None:
        self._next = 0

    def new(self) -> int:
        self._next += 1
        return self._next - 1

    def block(self, count: int) -> List[int]:
        return [self.new() for _ in range(count)]


def _text(*terms: Iterable[int]) -> List[str]:
    return ["".join(chr(97 + i) for i in term) for term in terms]


def _pos(state_objs: Sequence, wanted: Sequence) -> List[int]:
    """Positions (in state_objs) of ``wanted``, in the order wanted is given."""
    objs = list(state_objs)
    return [next(i for i, s in enumerate(objs) if s is w or s == w) for w in wanted]


def apply_operator_vector(state_objs: list, operator_objs: list) -> str:
    """``operator, state -> state`` on a (d0, ..., dn-1, 1) tensor, e.g. ``'eb,abcd->aecd'``
    (einsum_constructor.py:7-64)."""
    lab = _Labels()
    state = lab.block(len(state_objs))
    tail = lab.new()
    tgt = _pos(state_objs, operator_objs)
    fresh: Dict[int, int] = {t: lab.new() for t in tgt}
    op = [fresh[t] for t in tgt] + [state[t] for t in tgt]
    out = [fresh.get(i, state[i]) for i in range(len(state))] + [tail]
    a, b, c = _text(op, state + [tail], out)
    return f"{a},{b}->{c}"


def apply_operator_matrix(state_objs: list, operator_objs: list) -> str:
    """``operator, rho, conj(operator) -> rho`` on a (d..., d...) tensor
    (einsum_constructor.py:67-154)."""
    lab = _Labels()
    n = len(state_objs)
    rows, cols = lab.block(n), lab.block(n)
    tgt = _pos(state_objs, operator_objs)
    new_r = {t: lab.new() for t in tgt}
    left = [new_r[t] for t in tgt] + [rows[t] for t in tgt]
    new_c = {t: lab.new() for t in tgt}
    right = [new_c[t] for t in tgt] + [cols[t] for t in tgt]
    out = [new_r.get(i, rows[i]) for i in range(n)] + [new_c.get(i, cols[i]) for i in range(n)]
    a, b, c, d = _text(left, rows + cols, right, out)
    return f"{a},{b},{c}->{d}"


def trace_out_vector(state_objs: list, states: list) -> str:
    """Keep the labels of ``states`` (tensor order) and the trailing axis
    (einsum_constructor.py:157-192)."""
    lab = _Labels()
    state = lab.block(len(state_objs))
    tail = lab.new()
    keep = set(_pos(state_objs, states))
    a, b = _text(state + [tail], [state[i] for i in range(len(state)) if i in keep] + [tail])
    return f"{a}->{b}"


def trace_out_matrix(state_objs: list, states: list) -> str:
    """Partial trace, e.g. ``'abcadc->bd'`` keeps the middle one of three
    (einsum_constructor.py:195-230): a traced subsystem shares ONE label between its row and
    column axis, a kept one gets two fresh labels that also form the output."""
    lab = _Labels()
    kept = {getattr(s, "uid", id(s)) for s in states}
    shared: Dict[int, int] = {}
    operand: List[int] = []
    out: List[int] = []
    for _ in range(2):
        for i, so in enumerate(state_objs):
            if getattr(so, "uid", id(so)) in kept:
                c = lab.new()
                operand.append(c)
                out.append(c)
            else:
                if i not in shared:
                    shared[i] = lab.new()
                operand.append(shared[i])
    a, b = _text(operand, out)
    return f"{a}->{b}"


def reorder_vector(state_objs: list, states: list) -> str:
    """Transpose to the order of ``states`` (einsum_constructor.py:233-265); label 'a' is the
    trailing axis."""
    lab = _Labels()
    tail = lab.new()
    state = lab.block(len(state_objs))
    order = _pos(state_objs, states)
    a, b = _text(state + [tail], [state[i] for i in order] + [tail])
    return f"{a}->{b}"


def reorder_matrix(state_objs: list, states: list) -> str:
    """einsum_constructor.py:268-302."""
    lab = _Labels()
    n = len(state_objs)
    rows, cols = lab.block(n), lab.block(n)
    order = _pos(state_objs, states)
    a, b = _text(rows + cols, [rows[i] for i in order] + [cols[i] for i in order])
    return f"{a}->{b}"


def measure_vector(state_objs: list, states: list) -> str:
    """Same text as trace_out_vector (einsum_constructor.py:305-338)."""
    return trace_out_vector(state_objs, states)


def measure_matrix(state_objs: list, states: list) -> str:
    """Keep the row and column axes of ``states`` (einsum_constructor.py:341-370)."""
    lab = _Labels()
    n = len(state_objs)
    rows, cols = lab.block(n), lab.block(n)
    keep = set(_pos(state_objs, states))
    a, b = _text(rows + cols,
                 [rows[i] for i in range(n) if i in keep] + [cols[i] for i in range(n) if i in keep])
    return f"{a}->{b}"
