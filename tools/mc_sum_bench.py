#!/usr/bin/env python
"""Monte-Carlo sum exchange: the fused fold + NVLink peer-memory kernel (csrc/comm.cu) against the
two-kernel path (reduce_partials_kernel + NCCL all-reduce).  Launch with torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/mc_sum_bench.py

Reports (rank 0, one JSON line, device-timed with CUDA events, max over ranks):
  * latency of an all-reduce of n doubles alone, both ways, launched back to back inside one CUDA graph
  * cfg5 (one-loop numerator, Monte-Carlo sum) with the global sum produced EVERY step, both ways
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist

    import spenso_b200 as sp
    from spenso_b200 import workloads
    from spenso_b200.sharding import connect_peers

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    ctx = sp.Context(local, stream=stream.cuda_stream)
    comm = connect_peers(ctx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_graph(body, iters, reps=5):
        """median over reps of (one graph replay of `iters` bodies) / iters, in us, max over ranks"""
        with torch.cuda.stream(stream):
            for _ in range(3):
                body()
        barrier()
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g, stream=stream):
                for _ in range(iters):
                    body()
        except Exception as e:  # capture refused (e.g. a collective that cannot be captured): eager launches
            print(f"mc_sum_bench: graph capture failed ({type(e).__name__}); eager timing", file=sys.stderr)
            torch.cuda.synchronize()

            class Eager:
                def replay(self):
                    for _ in range(iters):
                        body()
            g = Eager()
        barrier()
        with torch.cuda.stream(stream):
            g.replay()
        barrier()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            with torch.cuda.stream(stream):
                e0.record(stream)
                g.replay()
                e1.record(stream)
            barrier()
            ts.append(e0.elapsed_time(e1) * 1e3 / iters)
        t = torch.tensor([float(np.median(ts))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    res = {"world": world, "allreduce_us": {}, "cfg5_sum_every_step": {}}
    # ---- the collective alone -------------------------------------------------------------------------------
    for n in (2, 32, 512):
        x = torch.ones(n, dtype=torch.float64, device=dev) * (rank + 1) * 1e-3
        peer = timed_graph(lambda: comm.allreduce_sum(x.data_ptr(), n), 200)
        x.fill_((rank + 1) * 1e-3)
        if world > 1:
            def nccl_body():
                dist.all_reduce(x)
                x.mul_(0.5)       # keeps the values bounded; one tiny elementwise launch rides along
            nccl = timed_graph(nccl_body, 200)
            def mul_body():
                x.mul_(0.5)
            mul_only = timed_graph(mul_body, 200)
            nccl = nccl - mul_only
        else:
            nccl = None
        res["allreduce_us"][str(n)] = {"peer_kernel": peer, "nccl": nccl}
    # ---- cfg5 with the global sum every step ------------------------------------------------------------------
    spec = workloads.one_loop_massive()
    n = 1 << 22
    arrs = workloads.synthetic_inputs(spec, n, 1237 + rank, scale=100.0, real=True)
    net = workloads.build_engine(spec, ctx)
    ins = []
    for node, a in zip([nd for nd in spec.nodes if nd[0] == "batched"], arrs):
        t = sp.BatchTensor.empty(ctx, node[1], n)
        t.upload(a)
        ins.append(t)
    size = net.result_size
    sums = torch.zeros(2 * size, dtype=torch.float64, device=dev)
    def local_step():
        net.execute(ins, None, sum_out_ptr=sums.data_ptr())

    def two_kernel():
        net.execute(ins, None, sum_out_ptr=sums.data_ptr())
        dist.all_reduce(sums)

    # the three variants are measured in alternation (the SM clock drifts under sustained FP64 load)
    t_local, t_fused, t_nccl = [], [], []
    fused_sum, agree = None, 0.0
    for _ in range(3):
        net.set_comm(None)
        t_local.append(timed_graph(local_step, 100, reps=3))
        net.set_comm(comm)
        t_fused.append(timed_graph(local_step, 100, reps=3))
        torch.cuda.synchronize()
        fused_sum = sums.clone()
        net.set_comm(None)
        if world > 1:
            t_nccl.append(timed_graph(two_kernel, 100, reps=3))
            torch.cuda.synchronize()
            agree = float((sums - fused_sum).abs().max() / fused_sum.abs().max())
    local_only, fused = float(np.median(t_local)), float(np.median(t_fused))
    nccl_step = float(np.median(t_nccl)) if t_nccl else None
    res["cfg5_sum_every_step"] = {
        "points_per_gpu_per_step": n, "us_per_step_local_sum_only": local_only,
        "us_per_step_fused_fold_exchange": fused, "us_per_step_fold_then_nccl": nccl_step,
        "evals_per_s_fused": world * n / (fused * 1e-6),
        "evals_per_s_nccl": (world * n / (nccl_step * 1e-6)) if nccl_step else None,
        "rel_diff_fused_vs_nccl": agree, "comm_status": comm.status,
    }
    barrier()
    if rank == 0:
        print(json.dumps(res), flush=True)
    comm.close(dist.barrier if world > 1 else None)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
