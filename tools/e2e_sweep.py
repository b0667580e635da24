import sys, time, numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import spenso_b200 as sp
from spenso_b200 import workloads
ctx = sp.Context(0)
spec = workloads.eemumu(); n = 1 << 20
net = workloads.build_engine(spec, ctx)
arrs = workloads.synthetic_inputs(spec, n, 1)
pinned = [torch.from_numpy(a.view(np.float64).reshape(n, -1)).pin_memory() for a in arrs]
host_in = [p.numpy().view(np.complex128).reshape(n, -1) for p in pinned]
out_t = torch.empty(n * 2, dtype=torch.float64).pin_memory()
out = out_t.numpy().view(np.complex128).reshape(n, 1)
for chunk in (1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20):
    for _ in range(3): net.execute_host(host_in, out=out, chunk_points=chunk)
    t0 = time.perf_counter()
    for _ in range(10): net.execute_host(host_in, out=out, chunk_points=chunk)
    dt = (time.perf_counter() - t0) / 10
    print(f"chunk {chunk:8d}: {dt*1e3:.3f} ms/step  {n/dt:.3e} evals/s  H2D {268.4/dt/1e3:.1f} GB/s", flush=True)
# plain single big copies for reference
d = [torch.empty_like(p, device="cuda") for p in pinned]
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    for a, b in zip(d, pinned): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"raw H2D of the 4 inputs: {dt*1e3:.3f} ms  {268.4/dt/1e3:.1f} GB/s")
