"""Eager stand-ins for the structured control flow helpers (TEST INFRASTRUCTURE ONLY)."""


def fori_loop(lower, upper, body, init):
    val = init
    for i in range(int(lower), int(upper)):
        val = body(i, val)
    return val


def while_loop(cond, body, init):
    val = init
    while bool(cond(val)):
        val = body(val)
    return val


def cond(pred, true_fn, false_fn, *operands, operand=None):
    if not operands:
        operands = (operand,)
    return true_fn(*operands) if bool(pred) else false_fn(*operands)
