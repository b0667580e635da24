"""Host-side logic of the N>1 path on CPU: world_size-2 gloo process group, point sharding and the final
all-reduce of Monte-Carlo partial sums (the only collective of the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spenso_b200.sharding import allreduce_partial_sums, shard_range


def test_shard_ranges_cover_exactly_once():
    for n in (0, 1, 7, 8, 1000, 1 << 20, 10 ** 8 + 3):
        for world in (1, 2, 3, 4, 8):
            nxt, total = 0, 0
            for r in range(world):
                first, cnt = shard_range(n, r, world)
                assert first == nxt and cnt >= 0
                nxt, total = first + cnt, total + cnt
            assert total == n
            sizes = [shard_range(n, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, n_points, out_size, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(42)                       # every rank sees the same global batch
    vals = rng.uniform(-1, 1, size=(n_points, out_size, 2))
    first, cnt = shard_range(n_points, rank, world)
    partial = torch.from_numpy(vals[first:first + cnt].sum(axis=0).reshape(-1).copy())   # stands in for the device partial
    allreduce_partial_sums(partial)
    q.put((rank, first, cnt, partial.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_gloo_allreduce_of_partial_sums():
    world, n_points, out_size = 2, 1001, 16
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_points, out_size, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    want = np.random.default_rng(42).uniform(-1, 1, size=(n_points, out_size, 2)).sum(axis=0).reshape(-1)
    covered = sorted((first, cnt) for _, first, cnt, _ in res)
    assert covered == [(0, 501), (501, 500)]
    for _, _, _, got in res:
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    assert (res[0][3] == res[1][3]).all()                 # every rank ends with the same estimate
