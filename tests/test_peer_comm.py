"""Monte-Carlo sum exchange over NVLink peer memory (csrc/comm.cu, spn_comm_*).

world = 1 runs on any GPU box (`-m gpu`); the world = 2 test spawns one process per GPU and is skipped on
boxes with a single GPU (run it with `gpurun --gpus 2`).  Host-side sharding logic: tests/test_sharding_gloo.py.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_comm_symbols_and_argument_checks():
    import spenso_b200 as sp
    from spenso_b200._capi import lib

    for name in ("spn_comm_create", "spn_comm_handle", "spn_comm_connect", "spn_comm_allreduce_sum", "spn_comm_status", "spn_comm_disconnect",
                 "spn_comm_destroy", "spn_net_set_comm"):
        assert hasattr(lib(), name)
    assert sp.PeerComm.HANDLE_BYTES == 64
    assert lib().spn_comm_status(None) == -1
    import ctypes as C
    out = C.c_void_p()
    assert lib().spn_comm_create(None, 0, 1, 16, C.byref(out)) != 0      # no context: refused, nothing allocated


@pytest.mark.gpu
def test_world_of_one_fold_and_exchange_matches_plain_sum():
    import torch
    import spenso_b200 as sp
    from spenso_b200 import workloads
    from spenso_b200.sharding import connect_peers

    ctx = sp.Context(0)
    comm = connect_peers(ctx)
    # plain all-reduce entry: identity for a world of one
    x = torch.arange(1, 33, dtype=torch.float64, device="cuda:0") * 0.37
    want = x.clone()
    comm.allreduce_sum(x.data_ptr(), 32)
    ctx.synchronize()
    assert torch.equal(x, want) and comm.status == 0
    # fused fold + exchange behind spn_net_execute: same bits as the two-kernel path
    spec = workloads.one_loop_massive()
    n = 50_000
    arrs = workloads.synthetic_inputs(spec, n, 7, scale=100.0, real=True)
    net = workloads.build_engine(spec, ctx)
    ins = []
    for node, a in zip(net_inputs(spec), arrs):
        t = sp.BatchTensor.empty(ctx, node, n)
        t.upload(a)
        ins.append(t)
    size = net.result_size
    s0 = torch.zeros(2 * size, dtype=torch.float64, device="cuda:0")
    s1 = torch.zeros(2 * size, dtype=torch.float64, device="cuda:0")
    net.execute(ins, None, sum_out_ptr=s0.data_ptr())
    net.set_comm(comm)
    launches = ctx.launch_count
    net.execute(ins, None, sum_out_ptr=s1.data_ptr())
    assert ctx.launch_count - launches == 2          # network kernel + ONE fold/exchange kernel
    ctx.synchronize()
    assert torch.equal(s0, s1) and float(s0.abs().max()) > 0
    # capturable: the epoch counter lives in device memory
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    ctx.set_stream(st.cuda_stream)
    st.wait_stream(torch.cuda.current_stream())
    s2 = torch.zeros_like(s1)
    with torch.cuda.graph(g, stream=st):
        for _ in range(3):
            net.execute(ins, None, sum_out_ptr=s2.data_ptr())
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(s0, s2) and comm.status == 0
    # an empty point range still runs the exchange (a rank whose shard is empty must not skip the collective)
    net.set_comm(comm)
    s3 = torch.full((2 * size,), 3.0, dtype=torch.float64, device="cuda:0")
    launches = ctx.launch_count
    net.execute(ins, None, sum_out_ptr=s3.data_ptr(), first_point=n)
    assert ctx.launch_count - launches == 1          # the exchange kernel alone
    ctx.synchronize()
    assert float(s3.abs().max()) == 0.0 and comm.status == 0
    # stepwise execution honours the communicator too
    step = workloads.build_engine(spec, ctx, sp.STEPWISE)
    s4 = torch.zeros_like(s0)
    step.set_comm(comm)
    launches = ctx.launch_count
    step.execute(ins, None, sum_out_ptr=s4.data_ptr())
    ctx.synchronize()
    assert float((s4 - s0).abs().max()) <= 1e-12 * float(s0.abs().max()) and comm.status == 0
    step.set_comm(None)
    s5 = torch.zeros_like(s0)
    l1 = ctx.launch_count
    step.execute(ins, None, sum_out_ptr=s5.data_ptr())
    assert (ctx.launch_count - l1) + 1 == l1 - launches      # the attached communicator added exactly one launch
    net.set_comm(None)
    comm.close()


def net_inputs(spec):
    """slot lists of the batched leaves, in input order"""
    return [node[1] for node in spec.nodes if node[0] == "batched"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), SPN_COMM_TIMEOUT_S="20")
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import spenso_b200 as sp
    from spenso_b200 import workloads
    from spenso_b200.sharding import connect_peers, shard_range

    try:
        torch.cuda.set_device(rank)
        dev = torch.device(f"cuda:{rank}")
        dist.init_process_group("nccl", device_id=dev)
        ctx = sp.Context(rank, stream=torch.cuda.current_stream().cuda_stream)
        comm = connect_peers(ctx)
        ok = True
        # (1) plain exchange vs NCCL, repeated (epoch parity), several sizes
        for it in range(40):
            n = (2, 32, 512)[it % 3]
            g = torch.Generator(device=dev).manual_seed(100 * it + rank)
            x = torch.rand(n, generator=g, device=dev, dtype=torch.float64) - 0.5
            y = x.clone()
            comm.allreduce_sum(x.data_ptr(), n)
            dist.all_reduce(y)
            torch.cuda.synchronize()
            gathered = [torch.empty_like(x) for _ in range(world)]
            dist.all_gather(gathered, x)
            ok &= all(torch.equal(gathered[0], t) for t in gathered)          # identical bits on every rank
            ok &= bool(torch.allclose(x, y, rtol=1e-15, atol=1e-15))
            if world == 2:
                ok &= bool(torch.equal(x, y))                                  # a + b is commutative
        # (2) sharded Monte-Carlo sum of cfg5 through the fused fold + exchange vs the one-rank sum
        spec = workloads.one_loop_massive()
        n_total = 200_003                                                       # ragged split
        arrs = workloads.synthetic_inputs(spec, n_total, 11, scale=100.0, real=True)
        first, cnt = shard_range(n_total, rank, world)
        net = workloads.build_engine(spec, ctx)
        size = net.result_size

        def run(lo, m, with_comm):
            ins = []
            for sl, a in zip([nd[1] for nd in spec.nodes if nd[0] == "batched"], arrs):
                t = sp.BatchTensor.empty(ctx, sl, max(m, 1))
                if m:
                    t.upload(a[lo:lo + m])
                ins.append(t)
            out = torch.zeros(2 * size, dtype=torch.float64, device=dev)
            net.set_comm(comm if with_comm else None)
            if m:
                net.execute(ins, None, sum_out_ptr=out.data_ptr(), n_points=m)
            elif with_comm:
                comm.allreduce_sum(out.data_ptr(), 2 * size)
            torch.cuda.synchronize()
            return out

        total = run(first, cnt, True)
        ref = run(0, n_total, False)
        scale = float(ref.abs().max())
        ok &= float((total - ref).abs().max()) <= 1e-12 * scale
        ok &= comm.status == 0
        # (3) more ranks than points: the ranks with an EMPTY shard still go through spn_net_execute (an empty point
        # range) and must take part in the exchange — fused and stepwise
        for mode in (sp.FUSED, sp.STEPWISE):
            netm = workloads.build_engine(spec, ctx, mode)
            netm.set_comm(comm)
            n_few = world - 1                       # the last rank owns nothing
            lo, m = shard_range(n_few, rank, world)
            ins = []
            for sl, a in zip([nd[1] for nd in spec.nodes if nd[0] == "batched"], arrs):
                t = sp.BatchTensor.empty(ctx, sl, max(m, 1))
                if m:
                    t.upload(a[lo:lo + m])
                ins.append(t)
            out = torch.full((2 * size,), 7.0, dtype=torch.float64, device=dev)
            if mode == sp.FUSED or m:
                netm.execute(ins, None, sum_out_ptr=out.data_ptr(), first_point=0 if m else 1)
            else:       # stepwise runs whole tensors: an empty shard contributes zeros through the plain exchange
                out.zero_()
                comm.allreduce_sum(out.data_ptr(), 2 * size)
            torch.cuda.synchronize()
            netm.set_comm(None)
            one = workloads.build_engine(spec, ctx, sp.FUSED)
            ins1 = []
            for sl, a in zip([nd[1] for nd in spec.nodes if nd[0] == "batched"], arrs):
                t = sp.BatchTensor.empty(ctx, sl, n_few)
                t.upload(a[:n_few])
                ins1.append(t)
            want = torch.zeros(2 * size, dtype=torch.float64, device=dev)
            one.execute(ins1, None, sum_out_ptr=want.data_ptr())
            torch.cuda.synchronize()
            ok &= float((out - want).abs().max()) <= 1e-12 * max(float(want.abs().max()), 1e-300)
            ok &= comm.status == 0
        net.set_comm(None)
        comm.close(dist.barrier)
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()))


@pytest.mark.gpu
def test_two_gpus_exchange_matches_nccl_and_single_rank_sum():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    if os.environ.get("TORCHELASTIC_RUN_ID") or os.environ.get("LOCAL_WORLD_SIZE"):
        # the test spawns one process per GPU itself: under torchrun every rank would spawn a full world of its own,
        # several ranks would share each GPU, and exchange kernels that wait for their peers must never time-share a GPU
        pytest.skip("run with plain pytest, not under torchrun")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in results), results
