"""CPU tier: the N>1 path on 2 ranks over gloo.  Each rank shoots its shard of the global packet
ids (host logic check `hc_shoot` of tests/hostcheck standing in for the kernel: same shoot_packet
code, same Philox streams); the exchange then follows the product's protocol (include/cmib.h
cmib_comm_exchange_and_update, there on NCCL): the accumulators are all-reduced, every rank updates
the chunks of cells it owns (cmib_owned_cell), equal-sized packs of the owned chunks are all-gathered.  The result must equal the single-rank run: identical counters, sums equal up to
summation order, and every rank ends with the same updated grid."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

PC = 3.086e16
NC = 16
NPK = 20001  # odd on purpose: the last rank takes the remainder


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shoot(lib_path, lo, cnt):
    hc = C.CDLL(lib_path)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ncells = NC ** 3
    rng = np.random.default_rng(3)
    cells = np.ascontiguousarray(np.stack([np.full(ncells, 1e8), np.exp(rng.uniform(np.log(1e-5), np.log(1e-2), ncells)),
                                           np.full(ncells, 1e-6), np.full(ncells, 8000.)], 1))
    acc = np.zeros(16 + ncells * 2)
    anchor = np.array([-5 * PC] * 3); sides = np.array([10 * PC] * 3)
    ncell = np.array([NC] * 3, np.int32); per = np.zeros(3, np.int32)
    ip = np.array([2, 0, 0, 2, 1, 0], np.int32)       # 2 sources, mono, FixedValue, FixedValue re-emission, H-only, Philox
    dp = np.array([(13.6 * 1.6021766208e-19) * (1 / 6.626070040e-34), 0., 0.364, 3.4e15])
    sp = np.array([0., 0., 0., PC, -PC, 0.5 * PC]); sw = np.array([0.3, 0.7])
    xs = np.zeros(14); xs[0] = 6.3e-22
    hc.hc_shoot(p(anchor), p(sides), p(ncell), p(per), p(cells), p(ip), p(dp), p(sp), p(sw), p(xs), C.c_uint64(cnt),
                C.c_uint64(lo), C.c_uint64(42), C.c_uint32(3), p(acc))
    return acc


def _update(J, heat):
    """stand-in for the per-cell state update: any deterministic function of the cell's sums"""
    return 1. / (1. + 1e-10 * J) + 1e-30 * heat


def _worker(rank, world, port, lib_path, out):
    import torch
    import torch.distributed as dist
    from cmacionize_b200 import capi
    from cmacionize_b200.distributed import shard_packets
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, cnt = shard_packets(NPK, rank, world)
    acc = torch.from_numpy(_shoot(lib_path, lo, cnt))
    ncells = NC ** 3
    # 1. the accumulators (counters + interleaved J, heat records) are summed on every rank
    dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    cells = acc[16:].view(ncells, 2)
    # 2. every rank updates the chunks it owns (chunk c of 1024 cells belongs to rank c % world)
    state = torch.full((ncells,), -1., dtype=torch.float64)
    n_owned = capi.owned_cell_count(ncells, world, rank)
    own = np.array([capi.owned_cell(j, world, rank) for j in range(n_owned)])
    state[own] = torch.from_numpy(_update(cells[own, 0].numpy(), cells[own, 1].numpy()))
    # 3. equal-sized packs of the owned chunks are all-gathered and scattered back into cell order
    n_pack = max(capi.owned_cell_count(ncells, world, r) for r in range(world))
    pack = torch.zeros(n_pack, dtype=torch.float64)
    pack[:n_owned] = state[own]
    packs = [torch.empty_like(pack) for _ in range(world)]
    dist.all_gather(packs, pack)
    for r in range(world):
        if r != rank:
            n_r = capi.owned_cell_count(ncells, world, r)
            state[np.array([capi.owned_cell(j, world, r) for j in range(n_r)])] = packs[r][:n_r]
    gathered = [torch.empty_like(cells) for _ in range(world)] if rank == 0 else None
    mine = cells.clone() if rank == 0 else torch.zeros_like(cells)   # the sums are the same on every rank
    dist.gather(mine, gathered, dst=0)
    states = [torch.empty_like(state) for _ in range(world)] if rank == 0 else None
    dist.gather(state, states, dst=0)
    if rank == 0:
        np.save(out, np.concatenate([acc[:16].numpy(), sum(g.numpy() for g in gathered).reshape(-1)]))
        np.save(out + ".state.npy", np.stack([t.numpy() for t in states]))
    dist.barrier()
    dist.destroy_process_group()


def test_distribute_matches_the_reference_formulas(cmib):
    """cmib_distribute / cmib_distribute_block = MPICommunicator::distribute / ::distribute_block
    (MPICommunicator.hpp:207-255): quotient (+1 for the first `remainder` ranks); blocks tile the range."""
    from cmacionize_b200 import capi
    for number in (0, 1, 7, 8, 9, 262144, 10**8, 16777216 + 5):
        for size in (1, 2, 3, 4, 8, 16):
            parts = [capi.distribute(number, size, r) for r in range(size)]
            q, rem = divmod(number, size)
            assert parts == [q + (1 if r < rem else 0) for r in range(size)] and sum(parts) == number
            blocks = [capi.distribute_block(r, size, 0, number) for r in range(size)]
            assert blocks[0][0] == 0 and blocks[-1][1] == number
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(size - 1))
            assert [b - a for a, b in blocks] == parts
    assert capi.distribute_block(1, 3, 100, 110) == (104, 107)


def test_owned_chunks_tile_the_grid(cmib):
    """cmib_owned_cell / cmib_owned_cell_count: chunks of 1024 cells dealt round-robin; every cell has one owner"""
    from cmacionize_b200 import capi
    for ncells in (1, 1023, 1024, 1025, 4096, 16 ** 3, 5000, 64 ** 3 + 7):
        for size in (1, 2, 3, 8):
            seen = np.zeros(ncells, int)
            for r in range(size):
                n = capi.owned_cell_count(ncells, size, r)
                cells = np.array([capi.owned_cell(j, size, r) for j in range(n)], dtype=np.int64)
                assert (cells < ncells).all() and (np.diff(cells) > 0).all()
                if size > 1:
                    assert ((cells // 1024) % size == r).all()
                seen[cells] += 1
            assert (seen == 1).all()


def test_two_rank_shoot_equals_single_rank(hostcheck, cmib, tmp_path):
    import torch.multiprocessing as mp
    from cmacionize_b200.distributed import shard_packets
    # shards tile the id range exactly
    for n, w in ((NPK, 2), (10, 3), (7, 8), (1000, 4)):
        sh = [shard_packets(n, r, w) for r in range(w)]
        assert sh[0][0] == 0 and sum(c for _, c in sh) == n
        assert all(sh[r][0] + sh[r][1] == sh[r + 1][0] for r in range(w - 1))
    lib_path = hostcheck._name
    single = _shoot(lib_path, 0, NPK)
    out = str(tmp_path / "acc.npy")
    mp.spawn(_worker, args=(2, _free_port(), lib_path, out), nprocs=2, join=True)
    both = np.load(out)
    states = np.load(out + ".state.npy")
    assert np.array_equal(states[0], states[1]) and (states[0] > 0.).all()   # every rank holds the whole updated grid
    cells = both[16:].reshape(-1, 2)
    assert np.array_equal(states[0], _update(cells[:, 0], cells[:, 1]))
    assert np.array_equal(both[:7], single[:7])          # weights by type, crossings, emissions: exact
    assert both[0] == NPK and both[6] > 1.05 * NPK       # re-emission happened
    scale = np.abs(single[16:]).max()
    assert np.abs(both[16:] - single[16:]).max() <= 1e-12 * scale
    assert abs(both[8] - single[8]) <= 1e-12 * single[8]   # optical depth traversed
