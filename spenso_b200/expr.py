"""Tiny expression builder for device-side leaf filling (spn_net_leaf_function).

    p = Param.vector(net_leaf_id, 4)          # the four components of a per-point momentum leaf
    e = sqrt(re(p[0] + p[3]))                 # ordinary Python arithmetic builds a tree ...
    code = compile_components([e, ...])       # ... that is flattened to the postfix program of the C ABI

Nothing is evaluated here: the program is inlined into the generated sm_100a kernel.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

PARAM, CONST, ADD, SUB, MUL, DIV, NEG, CONJ, SQRT, RE, IM, END = range(12)


class Expr:
    def __init__(self, op: int, args: Tuple["Expr", ...] = (), value=None):
        self.op, self.args, self.value = op, args, value

    @staticmethod
    def lift(x) -> "Expr":
        return x if isinstance(x, Expr) else Expr(CONST, value=complex(x))

    def __add__(self, o):
        return Expr(ADD, (self, Expr.lift(o)))

    __radd__ = lambda self, o: Expr.lift(o) + self

    def __sub__(self, o):
        return Expr(SUB, (self, Expr.lift(o)))

    __rsub__ = lambda self, o: Expr.lift(o) - self

    def __mul__(self, o):
        return Expr(MUL, (self, Expr.lift(o)))

    __rmul__ = lambda self, o: Expr.lift(o) * self

    def __truediv__(self, o):
        return Expr(DIV, (self, Expr.lift(o)))

    __rtruediv__ = lambda self, o: Expr.lift(o) / self

    def __neg__(self):
        return Expr(NEG, (self,))

    def conj(self):
        return Expr(CONJ, (self,))


class Param(Expr):
    """Component `flat` (canonical order) of the batched leaf `leaf` (a node id of the network)."""

    def __init__(self, leaf: int, flat: int):
        super().__init__(PARAM, value=(int(leaf), int(flat)))

    @staticmethod
    def vector(leaf: int, n: int) -> List["Param"]:
        return [Param(leaf, k) for k in range(n)]


def sqrt(x) -> Expr:
    """Square root of a real, non-negative quantity (wrap complex-typed inputs in re())."""
    return Expr(SQRT, (Expr.lift(x),))


def re(x) -> Expr:
    return Expr(RE, (Expr.lift(x),))


def im(x) -> Expr:
    return Expr(IM, (Expr.lift(x),))


def compile_components(components: Sequence) -> Tuple[List[Tuple[int, int]], List[complex], List[Tuple[int, int]]]:
    """-> (params, consts, code): the arrays spn_net_leaf_function takes; one program per component."""
    params: List[Tuple[int, int]] = []
    consts: List[complex] = []
    code: List[Tuple[int, int]] = []

    def emit(e: Expr):
        if e.op == PARAM:
            if e.value not in params:
                params.append(e.value)
            code.append((PARAM, params.index(e.value)))
        elif e.op == CONST:
            if e.value not in consts:
                consts.append(e.value)
            code.append((CONST, consts.index(e.value)))
        else:
            for a in e.args:
                emit(a)
            code.append((e.op, 0))

    for k, c in enumerate(components):
        if k:
            code.append((END, 0))
        emit(Expr.lift(c))
    return params, consts, code
