#!/usr/bin/env python3
"""How far does the state update (temperature solve) shrink with the number of cells?  lexingtonHII20 64^3 in its
steady state (1e7-packet iterations), then the update of contiguous blocks of 1/1, 1/2, 1/8, 1/64 of the cells around the
grid centre, timed with CUDA events; the accumulators are restored before every call so that each call solves the same
cells from the same state."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from cmacionize_b200 import problems

prob = problems.lexington(20, ncell=64, n_packets=10_000_000)
ctx = prob.ctx
for loop in range(8):
    problems.run_iteration(prob, loop)
n, T, x, _ = ctx.download_cells()
stream = torch.cuda.ExternalStream(ctx.stream())
nc = ctx.ncells
for rep in range(3):
    ctx.reset_accumulators()
    ctx.update_reemission_probabilities()
    ctx.shoot(10_000_000, seed=5, iteration=8 + rep)
    for frac in (1, 2, 8, 64):
        nb = nc // frac
        b0 = (nc - nb) // 2
        ctx.upload_cells(n, T, x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.update_state_block(8 + rep, 0., b0, b0 + nb)
        e1.record(stream)
        ctx.synchronize()
        print(f"rep {rep}: update of 1/{frac} of the cells ({nb}): {e0.elapsed_time(e1):.3f} ms")
ctx.close()
