#!/usr/bin/env python3
"""Small end-to-end runs for compute-sanitizer (memcheck / initcheck / racecheck / synccheck): every kernel
of an iteration on grids small enough for a 50x slowdown.  Queue capacity is forced small so that a shoot
takes several prepare -> march rounds with re-emission hand-over.  CMIB_SORT (0 / 2) selects the queue order.

  compute-sanitizer --tool memcheck python tools/sanitize_target.py
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ.setdefault("CMIB_QUEUE_CAPACITY", "16384")

from cmacionize_b200 import problems

# Lexington physics (full accumulator layout, Physical re-emission, temperature solve from loop 4)
prob = problems.lexington(20, ncell=12, n_packets=40000)
for loop in range(6):
    problems.run_iteration(prob, loop)
prob.ctx.synchronize()
T = prob.ctx.download_cells()[1] if hasattr(prob.ctx, "download_cells") else None
prob.ctx.close()
# Stromgren physics (H-only layout, no re-emission)
prob = problems.stromgren(ncell=12, n_packets=40000)
for loop in range(3):
    problems.run_iteration(prob, loop)
prob.ctx.synchronize()
prob.ctx.close()
print("sanitize target done", "" if T is None else f"T range {float(T.min()):.1f} .. {float(T.max()):.1f}")
