#!/usr/bin/env python3
"""Profile target: bring lexingtonHII20 64^3 to its ionised steady state, then run ONE shoot of
`--packets` packets between cudaProfilerStart/Stop, so that

  ncu --profile-from-start off --set full -k regex:'march_kernel|prepare_kernel' -c 6 python tools/profile_shoot.py

captures the first (full) rounds of the wavefront pipeline.  Also prints the shoot time measured
with CUDA events when run without a profiler."""
import argparse
import ctypes
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

ap = argparse.ArgumentParser()
ap.add_argument("--packets", type=float, default=4194304)
ap.add_argument("--problem", default="lexington", choices=["lexington", "stromgren", "stromgren256", "clumpy256", "clumpy256L"])
ap.add_argument("--algorithm", type=int, default=0)
ap.add_argument("--repeat", type=int, default=1)
ap.add_argument("--spinup-packets", type=float, default=2e6)
args = ap.parse_args()

import torch
from cmacionize_b200 import problems

n = int(args.packets)
if args.problem == "lexington":
    prob = problems.lexington(20, ncell=64, n_packets=n)
    spin = 7
elif args.problem == "stromgren":
    prob = problems.stromgren(ncell=64, n_packets=n)
    spin = 6
elif args.problem == "stromgren256":
    prob = problems.stromgren(ncell=256, n_packets=n)
    spin = 6
elif args.problem == "clumpy256L":
    prob = problems.synthetic_clumpy(ncell=256, n_packets=n, variant="Lexington")
    spin = 6
else:
    prob = problems.synthetic_clumpy(ncell=256, n_packets=n)
    spin = 6
ctx = prob.ctx
for loop in range(spin):
    problems.run_iteration(prob, loop, n_packets=int(args.spinup_packets))
ctx.synchronize()
ctx.set_shoot_algorithm(args.algorithm)
ctx.set_shoot_timing(True)
cudart = ctypes.CDLL("libcudart.so")
stream = torch.cuda.ExternalStream(ctx.stream())
for rep in range(args.repeat):
    ctx.reset_accumulators()
    ctx.update_reemission_probabilities()
    ctx.synchronize()
    cudart.cudaProfilerStart()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    ctx.shoot(n, seed=42, iteration=spin + rep, want_counters=False)
    e1.record(stream)
    ctx.synchronize()
    t1 = time.perf_counter()
    cudart.cudaProfilerStop()
    cross, emis = ctx.shoot_statistics()
    ms = e0.elapsed_time(e1)
    print(f"{args.problem}: shoot of {n} packets: {ms:.3f} ms (host {1e3*(t1-t0):.3f} ms) -> {n/ms*1e3:.3e} packets/s, "
          f"{cross/n:.2f} crossings/packet ({cross/ms*1e3:.3e} crossings/s), {emis/n:.3f} emissions/packet; "
          f"prepare {ctx.shoot_timing()[0]:.3f} ms, march {ctx.shoot_timing()[1]:.3f} ms, rounds {ctx.shoot_timing()[2]}, "
          f"lanes {ctx.shoot_overlap()[0]}, overlap {ctx.shoot_overlap()[1]:.3f} ms")
ctx.close()
