"""Achieved HBM bandwidth of the pairwise (Contract-trait) kernels on batched physics-sized tensors.
Usage (on a GPU box): python tools/pairwise_bench.py [n_points]
A 2^20-point launch of the small contractions lasts 65-95 us, where ramp-up and tail already cost several per cent (the
torch controls printed first show what plain streaming kernels of the same size reach on the same box); 2^22 points
show the kernels' steady state."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spenso_b200 as sp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
stream = torch.cuda.current_stream()
ctx = sp.Context(0, stream=stream.cuda_stream)
b = lambda a: sp.Bispinor.new_slot(4, a)
m = lambda a: sp.Minkowski.new_slot(4, a)


def rand_batch(sl, layout=sp.POINT_MAJOR):
    size = int(np.prod([s.dim for s in sl])) if sl else 1
    t = sp.BatchTensor.empty(ctx, sl, n, sp.C64, layout)
    x = torch.rand(n * size * 2, dtype=torch.float64, device="cuda")
    import ctypes
    # fill in place with a device-to-device copy (no host bounce)
    ctypes.CDLL("libcudart.so.12").cudaMemcpy(ctypes.c_void_p(t.device_ptr), ctypes.c_void_p(x.data_ptr()),
                                              ctypes.c_size_t(n * size * 16), 3)
    return t, size


def timeit(fn, iters=20):
    for _ in range(3):
        r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        r = fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, r


cases = []
A, sa = rand_batch([b(1), b(2), m(3)])
P, sp_ = rand_batch([m(3)])
cases.append(("dense[bis,bis,mink] x dense[mink] -> [bis,bis]   (K1)", lambda: A.contract(P), (sa + sp_ + 16) * 16))
M1, s1 = rand_batch([b(1), b(2)])
M2, s2 = rand_batch([b(2), b(3)])
cases.append(("dense[bis,bis] x dense[bis,bis] -> [bis,bis]      (K1)", lambda: M1.contract(M2), (s1 + s2 + 16) * 16))
ent = {(0, 2, 0): 1, (1, 3, 0): 1, (2, 0, 0): 1, (3, 1, 0): 1, (0, 3, 1): 1, (1, 2, 1): 1, (2, 1, 1): -1, (3, 0, 1): -1,
       (0, 3, 2): -1j, (1, 2, 2): 1j, (2, 1, 2): 1j, (3, 0, 2): -1j, (0, 2, 3): 1, (1, 3, 3): -1, (2, 0, 3): -1, (3, 1, 3): 1}
G = sp.SparseTensor.from_entries(ctx, [b(1), b(2), m(3)], ent)
cases.append(("sparse gamma (shared) x dense[mink] -> [bis,bis]   (K2)", lambda: G.contract(P), (sp_ + 16) * 16))
V, sv = rand_batch([b(2)])
cases.append(("dense[bis,bis] x dense[bis] -> [bis]              (K1)", lambda: M1.contract(V), (s1 + sv + 4) * 16))
cases.append(("dense[bis,bis] + dense[bis,bis]              (add)", lambda: M1 + M1, 3 * 16 * 16))
W, sw = rand_batch([b(1), b(2)])
cases.append(("dense[bis,bis] x dense[bis,bis] -> scalar (multi, K5)", lambda: M1.contract(W), (s1 + sw + 1) * 16))
M1c, _ = rand_batch([b(1), b(2)], sp.COMPONENT_MAJOR)
M2c, _ = rand_batch([b(2), b(3)], sp.COMPONENT_MAJOR)
cases.append(("component-major dense[bis,bis] x dense[bis,bis] -> [bis,bis]", lambda: M1c.contract(M2c), (s1 + s2 + 16) * 16))
Pc, _ = rand_batch([m(3)], sp.COMPONENT_MAJOR)
cases.append(("component-major sparse gamma x dense[mink] -> [bis,bis]", lambda: G.contract(Pc), (sp_ + 16) * 16))
# control: what this box's HBM gives a plain torch kernel of the same shape as the element-wise add (2 reads + 1 write)
xa = torch.rand(n * 32, dtype=torch.float64, device="cuda")
xb = torch.rand(n * 32, dtype=torch.float64, device="cuda")
xc = torch.empty_like(xa)
ms_ctl, _ = timeit(lambda: torch.add(xa, xb, out=xc))
print(f"{'control: torch.add of the same 768 B/point':62s} {ms_ctl*1e3:8.1f} us  {768*n/ms_ctl/1e6:7.0f} GB/s", flush=True)
ms_ctl, _ = timeit(lambda: xc.copy_(xa))
print(f"{'control: torch copy, 512 B/point':62s} {ms_ctl*1e3:8.1f} us  {512*n/ms_ctl/1e6:7.0f} GB/s", flush=True)
del xa, xb, xc
for name, fn, bytes_pp in cases:
    ms, r = timeit(fn)
    print(f"{name:62s} {ms*1e3:8.1f} us  {bytes_pp*n/ms/1e6:7.0f} GB/s  ({bytes_pp} B/point)", flush=True)
del r
# control: pure write and pure read streams on this box
xw = torch.empty(n * 32, dtype=torch.float64, device="cuda")
ms_ctl, _ = timeit(lambda: xw.fill_(1.0))
print(f"{'control: torch fill (write only), 256 B/point':62s} {ms_ctl*1e3:8.1f} us  {256*n/ms_ctl/1e6:7.0f} GB/s", flush=True)
ms_ctl, _ = timeit(lambda: torch.sum(xw))
print(f"{'control: torch sum (read only), 256 B/point':62s} {ms_ctl*1e3:8.1f} us  {256*n/ms_ctl/1e6:7.0f} GB/s", flush=True)
del xw
if os.environ.get("SPN_PWB_SMALL"):
    sys.exit(0)
# shapes outside the specialised-kernel thresholds: table-driven generic kernel / CSR gather
cases = []
co = lambda a: sp.ColorAdjoint.new_slot(8, a)
n_keep0, n = n, min(n, 1 << 18)
D3, sd3 = rand_batch([co(1), co(2), co(3)])
D2, sd2 = rand_batch([co(3), co(4)])
cases.append(("dense[coad^3] x dense[coad^2] -> [coad^3]  (row-parallel)", lambda: D3.contract(D2), (sd3 + sd2 + 512) * 16))
from spenso_b200 import workloads  # noqa: E402
F = sp.SparseTensor.from_entries(ctx, [co(1), co(2), co(4)], workloads.su3_f())
D3b, _ = rand_batch([co(1), co(2), co(3)])
cases.append(("dense[coad^3] . dense[coad^3] -> scalar  (streamed rows)", lambda: D3.contract(D3b), (2 * sd3 + 1) * 16))
cases.append(("dense[coad^3] x sparse f (shared) -> [coad^2]  (row-parallel)", lambda: D3.contract(F), (sd3 + 64) * 16))
for name, fn, bytes_pp in cases:
    ms, r = timeit(fn)
    print(f"{name:62s} {ms*1e3:8.1f} us  {bytes_pp*n/ms/1e6:7.0f} GB/s  ({bytes_pp} B/point)", flush=True)
del cases, r

# host cost of one a.contract(b) call (Python + ctypes + plan-cache lookup + allocation + launch): small batches, so
# that the device is never the bottleneck and the launch queue never fills
n_small = 4096
n_keep, n = n, n_small
a_s, _ = rand_batch([b(1), b(2)])
b_s, _ = rand_batch([b(2), b(3)])
for _ in range(200):
    r = a_s.contract(b_s)
torch.cuda.synchronize()
t0 = time.perf_counter()
N = 5000
for _ in range(N):
    r = a_s.contract(b_s)
t1 = time.perf_counter()
torch.cuda.synchronize()
import ctypes as C  # noqa: E402
from spenso_b200._capi import lib  # noqa: E402
h, m_, e_ = C.c_uint64(), C.c_uint64(), C.c_uint64()
lib().spn_ctx_plan_cache_stats(ctx.h, C.byref(h), C.byref(m_), C.byref(e_))
# the C call alone (no Python object construction around it)
out = C.c_void_p()
t2 = time.perf_counter()
for _ in range(N):
    lib().spn_tensor_contract(ctx.h, a_s.h, b_s.h, C.byref(out))
    lib().spn_tensor_free(out)
t3 = time.perf_counter()
torch.cuda.synchronize()
print(f"host cost per a.contract(b) call at {n_small} points: {(t1 - t0) / N * 1e6:.1f} us through the Python mirror, "
      f"{(t3 - t2) / N * 1e6:.1f} us for spn_tensor_contract + spn_tensor_free through ctypes; "
      f"plan cache: {h.value} hits, {m_.value} misses, {e_.value} entries", flush=True)
