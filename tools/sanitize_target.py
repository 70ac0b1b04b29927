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
# H-only layout with a heat term, fixed-probability re-emission, two periodic axes, two sources (the heat + periodic
# variants of march_lean_kernel under CMIB_SORT=2; the tail kernel in the H-only layout)
import numpy as np
from cmacionize_b200 import capi
ctx = capi.Context([-1e17, -2e17, -3e17], [2e17, 5e17, 3e17], [12, 20, 9], periodic=(True, False, True))
ctx.set_abundances()
sig = np.zeros(capi.NUM_IONS); sig[0] = 6.3e-22
ctx.set_cross_sections(capi.CROSS_SECTIONS_FIXED_VALUE, sig)
rr = np.zeros(capi.NUM_IONS); rr[0] = 4.e-19
ctx.set_recombination_rates(capi.RECOMBINATION_FIXED_VALUE, rr)
ctx.set_sources([[0., 0., 0.], [0.9e17, 2.9e17, -2.9e17]], [0.6, 0.4], 1e49)
ctx.set_spectrum(capi.SPECTRUM_MONOCHROMATIC, problems.ev_to_hz(15.))
ctx.set_reemission(capi.REEMISSION_FIXED_VALUE, 0.4, problems.ev_to_hz(14.2))
ctx.set_temperature_params(do_temperature_calculation=False)
nc = ctx.ncells
prob = problems.Problem("periodic_honly", ctx, np.full(nc, 3e8), np.full(nc, 8000.), problems.initial_fractions(nc), 40000, 1)
prob.upload()
for loop in range(3):
    problems.run_iteration(prob, loop)
ctx.synchronize()
ctx.close()
# the task-based packet conventions (abundance-weighted cross sections, unfolded in the update kernels)
prob = problems.lexington(20, ncell=10, n_packets=20000)
prob.ctx.set_packet_conventions(1)
for loop in range(6):
    problems.run_iteration(prob, loop)
prob.ctx.synchronize()
prob.ctx.close()
print("sanitize target done", "" if T is None else f"T range {float(T.min()):.1f} .. {float(T.max()):.1f}")
