"""Multi-GPU plumbing of the photoionization iteration: packets shard by global
id, the grid is replicated, ONE sum all-reduce per iteration combines the accumulator buffers
(counters ride in its first 8 doubles), every rank then runs the state update on all cells.

Replaces the reference's MPI path: `MPICommunicator::distribute` (src/MPICommunicator.hpp:207-222),
16 chunked `MPI_Allreduce` + 2 counter reductions (src/IonizationSimulation.cpp:410-416, 458-529)
and the 15 broadcast-based all-gathers after the state update (:540-618, not needed: the update is
replicated).  torch.distributed is the plumbing: NCCL on GPUs, gloo in the CPU-tier tests."""
from __future__ import annotations


def shard_packets(n_packets: int, rank: int, world: int):
    """(first global packet id, count) of this rank: contiguous blocks, remainder to the last rank.
    The packet's random stream depends only on (seed, iteration, global id), so the union over
    ranks is exactly the single-GPU packet set (MPICommunicator::distribute gives N/size (+1))."""
    per = n_packets // world
    lo = rank * per
    cnt = per if rank < world - 1 else n_packets - lo
    return lo, cnt


def accumulator_tensor(ctx, device):
    """zero-copy torch view of the context's accumulator buffer (counters + per-cell sums)"""
    import torch
    ptr, nd = ctx.accumulator_buffer()

    class _Buf:
        __cuda_array_interface__ = {"shape": (nd,), "typestr": "<f8", "data": (ptr, False), "version": 3}
    return torch.as_tensor(_Buf(), device=device)


def allreduce_sum(tensor):
    """the one collective of an iteration"""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor
