#!/usr/bin/env python3
"""Time the state update (temperature solve) of lexingtonHII20 64^3 in isolation: host clock around a
synchronised cmib_update_state, persistent kernel vs one-thread-per-cell kernel."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cmacionize_b200 import problems

npk = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
prob = problems.lexington(20, ncell=64, n_packets=npk)
ctx = prob.ctx
for loop in range(6):
    problems.run_iteration(prob, loop, n_packets=2_000_000)
for rep in range(3):
    for simple in ("0", "1"):
        os.environ["CMIB_UPDATE_SIMPLE"] = simple
        ctx.reset_accumulators()
        ctx.update_reemission_probabilities()
        ctx.shoot(npk, seed=42, iteration=6 + rep, want_counters=False)
        n0, T0, x0, _ = ctx.download_cells()
        ctx.synchronize()
        t0 = time.perf_counter()
        ctx.update_state(6 + rep, 0.)
        ctx.synchronize()
        t1 = time.perf_counter()
        print(f"rep {rep} simple={simple}: update_state {1e3*(t1-t0):.3f} ms")
        ctx.upload_cells(n0, T0, x0)
ctx.close()
