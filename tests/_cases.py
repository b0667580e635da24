"""Seeded random index structures and tensors shared by the host and GPU parity tests."""
import numpy as np

import _oracle as O

# (rep name, dim) pool — small dims keep tensors tiny; includes every rep kind the reference registers
POOL = [("euc", 2), ("euc", 3), ("bis", 4), ("bis", 2), ("coad", 3), ("mink", 4), ("mink", 2),
        ("lor", 3), ("lor", 2), ("spf", 2), ("cof", 3)]
DUALIZABLE = {"lor": "lor_d", "spf": "spf_d", "cof": "coaf"}


def random_pair(rng, max_free=3, max_common=3, allow_dual_first=True):
    """Two slot lists (argument order, shuffled) sharing `n_common` contracted indices."""
    n_common = int(rng.integers(0, max_common + 1))
    na = int(rng.integers(0, max_free + 1))
    nb = int(rng.integers(0, max_free + 1))
    aind = iter(rng.permutation(64).tolist())
    a, b = [], []
    for _ in range(n_common):
        rep, dim = POOL[int(rng.integers(len(POOL)))]
        i = next(aind)
        if rep in DUALIZABLE:
            if allow_dual_first and rng.random() < 0.5:
                a.append((DUALIZABLE[rep], dim, i)); b.append((rep, dim, i))
            else:
                a.append((rep, dim, i)); b.append((DUALIZABLE[rep], dim, i))
        else:
            a.append((rep, dim, i)); b.append((rep, dim, i))
    for lst, n in ((a, na), (b, nb)):
        for _ in range(n):
            rep, dim = POOL[int(rng.integers(len(POOL)))]
            if rep in DUALIZABLE and rng.random() < 0.5:
                rep = DUALIZABLE[rep]
            lst.append((rep, dim, next(aind)))
    a = [a[i] for i in rng.permutation(len(a))]
    b = [b[i] for i in rng.permutation(len(b))]
    return O.slots(*a) if a else np.zeros(0, O.SLOT_DTYPE), O.slots(*b) if b else np.zeros(0, O.SLOT_DTYPE)


def random_complex(rng, shape):
    return rng.uniform(-1, 1, size=shape) + 1j * rng.uniform(-1, 1, size=shape)


def random_sparse_entries(rng, dims, density=0.35):
    """{index tuple: value} with at least one entry (unless the tensor is empty)."""
    size = int(np.prod(dims)) if len(dims) else 1
    n = max(1, int(round(size * density)))
    flat = rng.choice(size, size=min(n, size), replace=False)
    out = {}
    for f in flat:
        idx = np.unravel_index(int(f), dims) if len(dims) else ()
        out[tuple(int(i) for i in idx)] = complex(rng.uniform(-1, 1), rng.uniform(-1, 1))
    return out
