#!/usr/bin/env python3
"""Wall time per iteration of the C++ host driver on the reference's benchmark files (as shipped:
stromgren 64^3, 1e6 packets x 20 iterations) — the small-N regime where launch overhead matters."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from cmacionize_b200 import host
name = sys.argv[1] if len(sys.argv) > 1 else "stromgren"
pf = ROOT / "tests" / "golden" / "benchmarks" / f"{name}.param"
sim = host.IonizationSimulation(pf)
sim.initialize()
for loop in range(3):
    sim.iteration(loop, sim.number_of_photons)
t0 = time.perf_counter()
n = 20
shoot = upd = 0.
for loop in range(3, 3 + n):
    r = sim.iteration(loop, sim.number_of_photons)
    shoot += r["shoot_s"]; upd += r["update_s"]
dt = time.perf_counter() - t0
print(f"{name}: {sim.number_of_photons} packets/iteration: {1e3*dt/n:.3f} ms per iteration "
      f"(shoot {1e3*shoot/n:.3f} ms, update {1e3*upd/n:.3f} ms) -> {sim.number_of_photons*n/dt:.3e} packets/s")
