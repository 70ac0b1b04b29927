"""GPU tier, 2 GPUs: the multi-GPU exchange of the C ABI (include/cmib.h cmib_comm_*: NCCL reduce of the
accumulators, state update of the owned cell chunks, all-gather of the opacity records)
under one process per GPU (torchrun, the launch bench.py uses) against the same iterations on one GPU."""
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SCRIPT = r"""
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from cmacionize_b200 import problems, capi
from cmacionize_b200.distributed import init_communicator, shard_packets, cell_block
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
which = sys.argv[2]
npk = 300001
def build():
    if which == "lexington":
        return problems.lexington(20, ncell=20, n_packets=npk, device=local)
    if which == "stromgren256":
        # H-only planes (the grid does not fit in L2), source at the threshold: the exchange sums the J_H plane only
        return problems.stromgren(ncell=256, n_packets=npk, device=local)
    return problems.stromgren(ncell=24, n_packets=npk, diffuse=True, device=local)
solo, team = build(), build()
init_communicator(team.ctx)
r, n, b0, b1 = team.ctx.comm_info()
assert (r, n) == (rank, world) and (b0, b1) == cell_block(team.ctx.ncells, rank, world), (r, n, b0, b1)
lo, cnt = shard_packets(npk, rank, world)
# iteration by iteration from the same state (tests/test_gpu_host_driver.py explains the tolerances); loop 5
# onwards solves the temperature (Lexington)
# (256^3 with 3e5 packets: a handful of packets per cell, so the order of the sums of iteration 0 shows more in iteration 1)
for loop, tol in ((0, 1e-10), (1, 1e-6 if which == "stromgren256" else 1e-7), (5, 1e-5 if which == "lexington" else 1e-9)):
    if loop == 5:
        # same starting state on both sides AND on every rank: the one-GPU runs of the two processes differ in the
        # last bits after two iterations (order of the atomic sums), and replicas that differ would shoot their
        # shares of the packets through different grids
        n0, T0, x0, _ = solo.ctx.download_cells()
        nc0 = n0.size
        t = torch.from_numpy(np.concatenate([n0, T0, np.nan_to_num(x0, nan=-1.).reshape(-1)])).cuda()
        dist.broadcast(t, src=0)
        t = t.cpu().numpy()
        n0, T0, x0 = t[:nc0].copy(), t[nc0:2 * nc0].copy(), t[2 * nc0:].reshape(14, nc0).copy()
        x0[x0 == -1.] = np.nan
        solo.ctx.upload_cells(n0, T0, x0)
        team.ctx.upload_cells(n0, T0, x0)
    for p, (plo, pcnt) in ((solo, (0, npk)), (team, (lo, cnt))):
        p.ctx.reset_accumulators()
        p.ctx.update_reemission_probabilities()
        tw, tc = p.ctx.shoot(pcnt, packet_offset=plo, seed=9, iteration=loop)
    solo.ctx.update_state(loop, 0.)
    team.ctx.exchange_and_update(loop, allreduce=(loop == 1))    # both reductions are exercised
    team.ctx.comm_gather_state()
    n1, T1, x1, h1 = solo.ctx.download_cells()
    n2, T2, x2, h2 = team.ctx.download_cells()
    ms = team.ctx.exchange_timing()
    assert all(m >= 0. for m in ms) and sum(ms) > 0., ms
    assert np.array_equal(n1, n2), 'densities differ'
    ok = np.isfinite(x1)
    assert np.array_equal(ok, np.isfinite(x2)), ('NaN pattern', loop, (~ok).sum(), (~np.isfinite(x2)).sum())
    assert np.abs(x2[ok] - x1[ok]).max() <= tol, (loop, np.abs(x2[ok] - x1[ok]).max())
    assert np.abs(T2 - T1).max() <= tol * 1e5, (loop, np.abs(T2 - T1).max())
    assert np.abs(h2 - h1).max() <= tol * max(np.abs(h1).max(), 1e-300)
    # every rank holds the same grid after the gathers: compare with rank 0 bit for bit
    t = torch.from_numpy(np.concatenate([T2, np.nan_to_num(x2).reshape(-1)])).cuda()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(t, ref), ('ranks hold different grids', loop, float((t - ref).abs().max()))
    # counters: every packet met the same fate as on one GPU
    c1, c2 = solo.ctx.shoot_statistics(), team.ctx.shoot_statistics()
    assert c1 == c2, (c1, c2)
# the distributed end-to-end form: upload the own block, gather, read the own block back
nc = team.ctx.ncells
n0, T0, x0, _ = solo.ctx.download_cells()
# the two processes' one-GPU runs differ in the last bits (order of the atomic sums): take rank 0's state everywhere
t = torch.from_numpy(np.concatenate([n0, T0, np.nan_to_num(x0).reshape(-1)])).cuda()
dist.broadcast(t, src=0)
t = t.cpu().numpy()
n0, T0, x0 = t[:nc].copy(), t[nc:2 * nc].copy(), t[2 * nc:].reshape(14, nc).copy()
team.ctx.upload_cells_block(b0, b1, np.ascontiguousarray(n0[b0:b1]), np.ascontiguousarray(T0[b0:b1] + 1.), np.ascontiguousarray(x0[:, b0:b1]))
team.ctx.comm_gather_cells_all()
n2, T2, x2, _ = team.ctx.download_cells()
assert np.array_equal(T2, T0 + 1.) and np.array_equal(np.nan_to_num(x2), np.nan_to_num(x0)), 'block upload + gather'
nb = b1 - b0
bn, bT, bx, bh = np.empty(nb), np.empty(nb), np.empty((14, nb)), np.empty((2, nb))
team.ctx.download_cells_block_into(b0, b1, bn, bT, bx, bh)
assert np.array_equal(bT, T0[b0:b1] + 1.) and np.array_equal(bn, n0[b0:b1]), 'block download'
# the distributed read-back: the cells the rank updated, in work-item order, against the full download
no = team.ctx.owned_cells()
assert no == capi.owned_cell_count(nc, world, rank)
own = np.array([capi.owned_cell(j, world, rank) for j in range(no)])
on, oT, ox, oh = np.empty(no), np.empty(no), np.empty((14, no)), np.empty((2, no))
_, _, _, h2 = team.ctx.download_cells()
team.ctx.download_cells_owned_into(on, oT, ox, oh)
assert np.array_equal(oT, T2[own]) and np.array_equal(on, n2[own]) and np.array_equal(np.nan_to_num(ox), np.nan_to_num(x2[:, own]))
assert np.array_equal(oh, h2[:, own]), 'owned download'
# ... and the distributed upload: every rank uploads the cells it owns, the opacity records are gathered
team.ctx.upload_cells_owned(on, oT + 2., ox)
team.ctx.comm_gather_owned_cells()
n3, T3, x3, _ = team.ctx.download_cells()
assert np.array_equal(T3, T2 + 2.) and np.array_equal(n3, n2) and np.array_equal(np.nan_to_num(x3[:2]), np.nan_to_num(x2[:2])), 'owned upload + gather'
assert np.array_equal(np.nan_to_num(x3[2:, own]), np.nan_to_num(x2[2:, own]))
team.ctx.comm_finalize()
dist.barrier()
if rank == 0:
    print("COMM-OK")
dist.destroy_process_group()
"""


def _ngpu():
    out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
    return len([l for l in out.splitlines() if l.startswith("GPU ")])


@pytest.mark.parametrize("which", ["stromgren_diffuse", "lexington", "stromgren256"])
def test_two_rank_exchange_equals_one_gpu(tmp_path, which):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    script = tmp_path / "comm.py"
    script.write_text(SCRIPT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", str(script), str(ROOT), which],
                         capture_output=True, text=True, timeout=600)
    if out.returncode != 0:
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / f"comm_test_{which}.err").write_text(out.stderr)
    assert out.returncode == 0 and "COMM-OK" in out.stdout, (out.stdout[-1500:], out.stderr[-6000:])


def test_scatter_rates_are_measured(cmib):
    """cmib_measure_scatter_rates: the roofline denominators bench.py measures in its own run"""
    with cmib.Context([0, 0, 0], [1, 1, 1], [32, 32, 32]) as ctx:
        red, gather = ctx.measure_scatter_rates(64 ** 3)
    assert 1e10 < red < 1e12 and 1e10 < gather < 1e12
