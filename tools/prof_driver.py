#!/usr/bin/env python
"""Launch ONE of the engine's kernels a few times — the command `ncu` wraps (tools/profile_r02.sh).
usage: prof_driver.py cfg2|cfg5|cfg5r|cfg3p|cfg3c|pair1|pair3|rows|zgemm|dgemm|mixed|exchange [n_points]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spenso_b200 as sp  # noqa: E402
from spenso_b200 import workloads  # noqa: E402

what = sys.argv[1]
stream = torch.cuda.current_stream()
ctx = sp.Context(0, stream=stream.cuda_stream)
dev = torch.device("cuda:0")
ITERS = 6


def network(wl, n, layout, want_sum):
    spec = workloads.WORKLOADS[wl]()
    net = workloads.build_engine(spec, ctx, sp.FUSED)
    g = torch.Generator(device=dev).manual_seed(1)
    ins, keep = [], []
    for node, (_, size), dt in zip([x for x in spec.nodes if x[0] == "batched"], spec.inputs, spec.input_dtypes):
        w = 2 if dt == sp.C64 else 1
        shape = (n, size, w) if layout == sp.POINT_MAJOR else (size, n, w)
        x = (torch.rand(shape, generator=g, device=dev, dtype=torch.float64) * 2 - 1) * (100.0 if wl.startswith("cfg5") else 1.0)
        if wl.startswith("cfg5") and dt == sp.C64:
            x[..., 1] = 0
        keep.append(x)
        ins.append(sp.BatchTensor.wrap(ctx, node[1], n, x.data_ptr(), dtype=dt, layout=layout, keepalive=x))
    out = None if want_sum else sp.BatchTensor.empty(ctx, net.result_slots, n, sp.C64, layout)
    sums = torch.zeros(2 * net.result_size, dtype=torch.float64, device=dev)
    for _ in range(ITERS):
        net.execute(ins, out, sum_out_ptr=sums.data_ptr() if want_sum else 0)
    torch.cuda.synchronize()


def rand_batch(sl, n, dtype=sp.C64):
    size = int(np.prod([s.dim for s in sl])) if sl else 1
    w = 2 if dtype == sp.C64 else 1
    x = torch.rand(n * size * w, dtype=torch.float64, device=dev) * 2 - 1
    return sp.BatchTensor.wrap(ctx, sp.canonicalize(sl)[0], n, x.data_ptr(), dtype=dtype, keepalive=x)


b = lambda a: sp.Bispinor.new_slot(4, a)
m = lambda a: sp.Minkowski.new_slot(4, a)
e = lambda d, a: sp.Euclidean.new_slot(d, a)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if what == "cfg2":
    network("cfg2", n or 1 << 20, sp.POINT_MAJOR, False)
elif what in ("cfg5", "cfg5r"):
    network(what, n or 1 << 22, sp.POINT_MAJOR, True)
elif what == "cfg3p":
    network("cfg3", n or 1 << 20, sp.POINT_MAJOR, False)
elif what == "cfg3c":
    network("cfg3", n or 1 << 20, sp.COMPONENT_MAJOR, False)
elif what in ("pair1", "pair3"):
    n = n or 1 << 20
    if what == "pair1":
        A, B = rand_batch([b(1), b(2), m(3)], n), rand_batch([m(3)], n)
    else:
        ent = {(0, 2, 0): 1, (1, 3, 0): 1, (2, 0, 0): 1, (3, 1, 0): 1, (0, 3, 1): 1, (1, 2, 1): 1, (2, 1, 1): -1, (3, 0, 1): -1,
               (0, 3, 2): -1j, (1, 2, 2): 1j, (2, 1, 2): 1j, (3, 0, 2): -1j, (0, 2, 3): 1, (1, 3, 3): -1, (2, 0, 3): -1, (3, 1, 3): 1}
        A, B = sp.SparseTensor.from_entries(ctx, [b(1), b(2), m(3)], ent), rand_batch([m(3)], n)
    for _ in range(ITERS):
        r = A.contract(B)
    torch.cuda.synchronize()
elif what == "rows":
    n = n or 1 << 18
    c = lambda a: sp.ColorAdjoint.new_slot(8, a)
    A, B = rand_batch([c(1), c(2), c(3)], n), rand_batch([c(3), c(4)], n)
    for _ in range(ITERS):
        r = A.contract(B)
    torch.cuda.synchronize()
elif what in ("zgemm", "dgemm", "mixed"):
    d = n or 16
    da = sp.F64 if what in ("dgemm", "mixed") else sp.C64
    db = sp.F64 if what == "dgemm" else sp.C64
    A = rand_batch([e(d, i) for i in (1, 2, 3, 4, 5, 6)], 1, da)
    B = rand_batch([e(d, i) for i in (4, 5, 6, 7, 8, 9)], 1, db)
    for _ in range(4):
        r = A.contract(B)
    torch.cuda.synchronize()
elif what == "exchange":
    from spenso_b200.sharding import connect_peers

    comm = connect_peers(ctx)
    x = torch.ones(32, dtype=torch.float64, device=dev)
    for _ in range(ITERS):
        comm.allreduce_sum(x.data_ptr(), 32)
    torch.cuda.synchronize()
    comm.close()
else:
    raise SystemExit(f"unknown target {what}")
print("ok", what)
