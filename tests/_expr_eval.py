"""numpy evaluation of spenso_b200.expr trees — TEST infrastructure (the product inlines these expressions into
its CUDA kernel and never evaluates them on the host)."""
import numpy as np

from spenso_b200 import expr as X


def evaluate(e, leaf_arrays):
    """leaf_arrays: {leaf id: complex [n_points, size]} -> complex [n_points]."""
    e = X.Expr.lift(e)
    if e.op == X.PARAM:
        leaf, flat = e.value
        return leaf_arrays[leaf][:, flat].astype(np.complex128)
    if e.op == X.CONST:
        n = next(iter(leaf_arrays.values())).shape[0]
        return np.full(n, e.value, dtype=np.complex128)
    a = [evaluate(x, leaf_arrays) for x in e.args]
    if e.op == X.ADD:
        return a[0] + a[1]
    if e.op == X.SUB:
        return a[0] - a[1]
    if e.op == X.MUL:
        return a[0] * a[1]
    if e.op == X.DIV:
        return a[0] / a[1]
    if e.op == X.NEG:
        return -a[0]
    if e.op == X.CONJ:
        return a[0].conj()
    if e.op == X.SQRT:
        assert not a[0].imag.any()
        return np.sqrt(a[0].real).astype(np.complex128)
    if e.op == X.RE:
        return a[0].real.astype(np.complex128)
    if e.op == X.IM:
        return a[0].imag.astype(np.complex128)
    raise ValueError(e.op)
