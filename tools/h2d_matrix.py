#!/usr/bin/env python
"""What the host gives: concurrent host<->device copy rates of this box when 1..N ranks copy at the same time.

Launch with torchrun (one rank per GPU), e.g. on an 8-GPU box
  for n in 1 2 4 8; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port 29533 tools/h2d_matrix.py; done
Rank 0 prints one JSON line per world size: GB/s PER GPU (max-over-ranks time) for
  * source memory: pinned (cudaHostAllocDefault) and write-combined pinned (spn_host_alloc_flags)
  * direction: H2D alone, D2H alone, both at once (the e2e leg of bench.py moves 256 MiB in + 16 MiB out per step)
  * 1 or 2 H2D streams per GPU, chunk sizes 8 / 32 / 256 MiB
This is the ceiling any host-fed throughput of bench.py can reach at that N (e2e.link_ceiling)."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist

    from spenso_b200._capi import lib

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    total = 256 << 20                       # bytes per direction per measurement (the H2D volume of one cfg2 step)
    dst = torch.empty(total // 8, dtype=torch.float64, device=dev)
    back = torch.empty(total // 8, dtype=torch.float64, device=dev)
    pinned = torch.empty(total // 8, dtype=torch.float64).pin_memory()
    pinned.uniform_(-1, 1)
    pinned_out = torch.empty(total // 8, dtype=torch.float64).pin_memory()
    wc_ptr = lib().spn_host_alloc_flags(total, 1)
    wc = None
    if wc_ptr:
        buf = (C.c_double * (total // 8)).from_address(wc_ptr)
        wc_np = np.frombuffer(buf, dtype=np.float64)
        wc_np[:] = pinned.numpy()            # sequential CPU writes: what write-combined memory is for
        wc = torch.from_numpy(wc_np)
    cudart = C.CDLL("libcudart.so.12")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rate(src, h2d, d2h, streams, chunk_mb, d2h_bytes=total):
        """GB/s per GPU of the H2D direction (or D2H when h2d is False), all ranks at once, best of 3"""
        chunk = chunk_mb << 20
        ss = [torch.cuda.Stream(device=dev) for _ in range(streams)]
        so = torch.cuda.Stream(device=dev)
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            cur = torch.cuda.current_stream()
            e0.record(cur)
            for s in ss + [so]:
                s.wait_stream(cur)
            if h2d:
                for k, lo in enumerate(range(0, total, chunk)):
                    n = min(chunk, total - lo)
                    st = ss[k % streams]
                    cudart.cudaMemcpyAsync(C.c_void_p(dst.data_ptr() + lo), C.c_void_p(src.data_ptr() + lo), C.c_size_t(n), 1,
                                           C.c_void_p(st.cuda_stream))
            if d2h:
                for lo in range(0, d2h_bytes, chunk):
                    n = min(chunk, d2h_bytes - lo)
                    cudart.cudaMemcpyAsync(C.c_void_p(pinned_out.data_ptr() + lo), C.c_void_p(back.data_ptr() + lo), C.c_size_t(n), 2,
                                           C.c_void_p(so.cuda_stream))
            for s in ss + [so]:
                cur.wait_stream(s)
            e1.record(cur)
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        moved = total if h2d else d2h_bytes
        return moved / (best * 1e-3) / 1e9

    res = {"world": world, "bytes_per_direction": total, "gbs_per_gpu": {}}
    for name, src in (("pinned", pinned), ("write_combined", wc)):
        if src is None:
            continue
        r = {}
        for streams in (1, 2):
            for chunk_mb in (8, 32, 256):
                r[f"h2d_{streams}stream_{chunk_mb}MiB"] = rate(src, True, False, streams, chunk_mb)
        r["h2d_with_d2h_16MiB_1stream_32MiB"] = rate(src, True, True, 1, 32, d2h_bytes=16 << 20)
        r["h2d_with_d2h_256MiB_1stream_32MiB"] = rate(src, True, True, 1, 32)
        res["gbs_per_gpu"][name] = r
    res["gbs_per_gpu"]["d2h_alone_32MiB"] = rate(pinned, False, True, 1, 32)
    barrier()
    if rank == 0:
        print(json.dumps(res), flush=True)
    if wc_ptr:
        del wc
        lib().spn_host_free(C.c_void_p(wc_ptr))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
