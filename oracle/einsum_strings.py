"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restatement of the index semantics the reference generates in
``photon_weave/extra/einsum_constructor.py``.  The reference emits strings with
labels ``chr(97 + i)``; here every generator returns *integer sublists* (usable
with ``np.einsum(op, sub, ...)`` beyond 26 labels) and ``to_string`` renders the
reference's exact string for the golden checks
(``tests/utils/test_einsum_constructor.py:15-16`` expects ``'abcadc->bd'``).

Positions (ints) stand in for the reference's state objects.
"""

from __future__ import annotations

import itertools
from typing import List, Sequence, Tuple


def to_string(operands: Sequence[Sequence[int]], out: Sequence[int]) -> str:
    """Render sublists with the reference's labelling rule (chr(97+i))."""
    lhs = ",".join("".join(chr(97 + i) for i in s) for s in operands)
    return f"{lhs}->{''.join(chr(97 + i) for i in out)}"


def apply_operator_vector(
    n_states: int, targets: Sequence[int]
) -> Tuple[List[List[int]], List[int]]:
    """
    einsum_constructor.py:7-64.  Operands: (operator, state(+trailing 1 axis)).
    State labels are allocated first (0..n-1), then the trailing axis, then one
    new label per target *in the order the targets are given*; the operator is
    indexed (new_targets..., old_targets...).
    """
    counter = itertools.count()
    state = [next(counter) for _ in range(n_states)]
    extra = next(counter)
    new = {t: next(counter) for t in targets}
    op = [new[t] for t in targets] + [state[t] for t in targets]
    out = [new[i] if i in new else state[i] for i in range(n_states)] + [extra]
    return [op, state + [extra]], out


def apply_operator_matrix(
    n_states: int, targets: Sequence[int]
) -> Tuple[List[List[int]], List[int]]:
    """
    einsum_constructor.py:67-154.  Operands: (operator, rho, conj(operator)).
    rho labels: rows 0..n-1, columns n..2n-1.  Left operator: (new_row..., row...),
    right operator: (new_col..., col...).
    """
    counter = itertools.count()
    rows = [next(counter) for _ in range(n_states)]
    cols = [next(counter) for _ in range(n_states)]
    new_r = {t: next(counter) for t in targets}
    op_l = [new_r[t] for t in targets] + [rows[t] for t in targets]
    new_c = {t: next(counter) for t in targets}
    op_r = [new_c[t] for t in targets] + [cols[t] for t in targets]
    out = [new_r[i] if i in new_r else rows[i] for i in range(n_states)] + [
        new_c[i] if i in new_c else cols[i] for i in range(n_states)
    ]
    return [op_l, rows + cols, op_r], out


def kraus_matrix(
    n_states: int, targets: Sequence[int]
) -> Tuple[List[List[int]], List[int]]:
    """
    core/adapters.py:170-207 (_cached_kraus_einsum): the matrix string with one
    extra label shared by both operator operands (summed: it is absent from the
    output).  NOTE the reference docstring shows '->kbd' but the code builds
    '->' + rhs of the base string, i.e. the Kraus label IS summed.
    """
    operands, out = apply_operator_matrix(n_states, targets)
    used = {i for s in operands for i in s} | set(out)
    k = ord("k") - 97  # the reference starts at 'k' and skips labels in use
    while k in used:
        k += 1
    operands[0] = [k] + operands[0]
    operands[2] = [k] + operands[2]
    return operands, out


def trace_out_matrix(
    n_states: int, keep: Sequence[int]
) -> Tuple[List[List[int]], List[int]]:
    """
    einsum_constructor.py:195-230.  Kept subsystems get fresh labels on the row
    and on the column side (and appear in the output in TENSOR order, rows then
    columns); traced subsystems share one label between row and column.
    """
    counter = itertools.count()
    shared = {}
    operand: List[int] = []
    out: List[int] = []
    for _ in range(2):
        for i in range(n_states):
            if i not in keep:
                if i not in shared:
                    shared[i] = next(counter)
                operand.append(shared[i])
            else:
                c = next(counter)
                operand.append(c)
                out.append(c)
    return [operand], out
