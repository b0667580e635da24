"""Multi-GPU plumbing: phase-space points are independent, so the batch is cut into contiguous blocks,
one per rank; nothing is exchanged while the network is evaluated.  The only collective
of the path is the all-reduce of the Monte-Carlo partial sums (2*|out| doubles): connect_peers() gives the
library's own exchange over NVLink peer memory (csrc/comm.cu), allreduce_partial_sums() the torch.distributed
one (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_points: int, rank: int, world: int) -> Tuple[int, int]:
    """(first_point, count) of rank's block: blocks are contiguous, ordered by rank, cover [0, n_points)
    exactly once and differ in size by at most one point (ragged totals allowed, empty blocks allowed)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_points, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def allreduce_partial_sums(partial, group=None):
    """In-place SUM all-reduce of a tensor of per-rank Monte-Carlo partial sums (torch tensor, f64;
    NCCL over NVLink on GPUs, gloo in the CPU tests).  No-op without an initialised process group."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def connect_peers(ctx, group=None, max_f64: int = 512):
    """PeerComm over the ranks of the (initialised) torch.distributed group: the 64-byte mailbox handles are
    all-gathered through torch.distributed; the exchange itself is csrc/comm.cu, not NCCL.
    Without a process group: a world of one."""
    import torch
    import torch.distributed as dist

    from . import PeerComm

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return PeerComm(ctx, 0, 1, max_f64=max_f64)
    rank, world = dist.get_rank(group), dist.get_world_size(group)

    def exchange(mine: bytes):
        dev = f"cuda:{ctx.device}" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t, group=group)
        return [bytes(o.cpu().tolist()) for o in out]

    comm = PeerComm(ctx, rank, world, exchange, max_f64=max_f64)
    dist.barrier(group)      # every rank has mapped every mailbox before the first exchange
    return comm
