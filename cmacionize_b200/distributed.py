"""Multi-GPU plumbing of the photoionization iteration for one-process-per-GPU launches
(torchrun): the collectives themselves live behind the C ABI (include/cmib.h `cmib_comm_*`: NCCL on the
context's stream — reduce-scatter of the accumulators, block update, gather of the opacity records), exactly
the code the C++ driver (`IonizationSimulation(..., devices)`, one host thread per GPU) runs.  What is left
here is the rendezvous: rank 0 creates the NCCL unique id and torch.distributed carries its 128 bytes to the
other ranks.

Replaces the reference's MPI path: `MPICommunicator::distribute` / `distribute_block`
(src/MPICommunicator.hpp:207-255), 16 chunked `MPI_Allreduce` + 2 counter reductions
(src/IonizationSimulation.cpp:410-416, 458-529) and the 15 broadcast-based all-gathers after the block-wise
state update (:540-618)."""
from __future__ import annotations

from . import capi


def shard_packets(n_packets: int, rank: int, world: int):
    """(first global packet id, count) of this rank: the contiguous id blocks of
    MPICommunicator::distribute_block, whose sizes are those of ::distribute (quotient, +1 for the first
    `remainder` ranks).  A packet's random stream depends only on (seed, iteration, global id), so the union
    over the ranks is exactly the single-GPU packet set."""
    lo, hi = capi.distribute_block(rank, world, 0, n_packets)
    return lo, hi - lo


def cell_block(n_cells: int, rank: int, world: int):
    """[first, one past last) cell this rank updates (MPICommunicator::distribute_block)"""
    return capi.distribute_block(rank, world, 0, n_cells)


def init_communicator(ctx, rank: int | None = None, world: int | None = None):
    """Give `ctx` its rank of a communicator that spans the torch.distributed world (collective).  The id
    travels as a CPU byte tensor through the default process group (gloo or NCCL)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    ids = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init_rank(world, rank, ids[0])
