#!/usr/bin/env python3
"""bench.py — benchmark of the photoionization hot path.

Headline metric (BASELINE.json): photon packets/s on lexingtonHII20 and wall time per ionization iteration.
A *step* is one full iteration of IonizationSimulation::run's loop body (reference
src/IonizationSimulation.cpp:359-643) on the Lexington HII20 benchmark (64^3 cells, 1e8 packets per iteration,
Planck 20 000 K, Verner cross sections, 14 ions + 2 heating terms, Physical diffuse re-emission, temperature solve
with line cooling):

    reset accumulators -> re-emission probabilities -> shoot -> exchange (N > 1) -> state update

After the headline the same loop is timed on the two north-star grids (BASELINE.json configs[4] and the 256^3
Stroemgren grid of the target) at the 1e9 packets per iteration configs[4] names, and attached under
`workloads.{stromgren256, clumpy256}`; each entry carries its own `value`, `ms_per_step`, `roofline` (HBM: these
grids do not fit in L2), `phases_ms` and `clocks`.

Usage:  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workloads a,b,...]
N > 1 is launched by torch.distributed.run (one rank per GPU).  Scaling is STRONG, as in the reference
(IonizationSimulation.cpp:392-397 -> MPICommunicator::distribute): the iteration keeps the configuration's packet
count and the ranks split it by global packet id; the grid is replicated; the exchange between shoot and update is
the product's own NCCL code behind the C ABI (include/cmib.h cmib_comm_exchange_and_update: the accumulators are
all-reduced, every rank updates the cell chunks it owns, the opacity records are all-gathered) — the
same calls the C++ driver `CMacIonizeB200 --gpus` makes.  `weak` (N > 1) adds the figure with N x the packets.

One JSON line is printed by rank 0 (schema: task contract + `roofline`, `cpu_baseline`, `e2e`, `clocks`,
`gpu_launches`, `phases_ms`, `workloads`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "photon packets/s on lexingtonHII20 (whole iteration: shoot + state update)"
UNIT = "packets/s"
FULL_PACKETS = 100_000_000

LEXINGTON_PARAM = """AbundanceModel:
  type: FixedValue
  He: 0.1
  C: 2.2e-4
  N: 4.e-5
  O: 3.3e-4
  Ne: 5.e-5
  S: 9.e-6
ContinuousPhotonSource:
  type: None
DensityFunction:
  type: BlockSyntax
  filename: {yml}
DiffuseReemissionHandler:
  type: Physical
SimulationBox:
  anchor: [-3. pc, -3. pc, -3. pc]
  sides: [6. pc, 6. pc, 6. pc]
  periodicity: [false, false, false]
DensityGrid:
  type: Cartesian
  number of cells: [{nc}, {nc}, {nc}]
IonizationSimulation:
  output folder: .
  number of iterations: {nit}
  number of photons: {npk}
  random seed: 42
TemperatureCalculator:
  do temperature calculation: true
  PAH heating factor: 0.
PhotonSourceDistribution:
  type: SingleStar
  position: [0. pc, 0. pc, 0. pc]
  luminosity: 1.e49 s^-1
PhotonSourceSpectrum:
  type: Planck
  temperature: 20000. K
"""
LEXINGTON_YML = """number of blocks: 2
block[0]:
  origin: [0. pc, 0. pc, 0. pc]
  sides: [6. pc, 6. pc, 6. pc]
  type: cube
  number density: 100. cm^-3
  initial temperature: 8000. K
block[1]:
  origin: [0. pc, 0. pc, 0. pc]
  sides: [6.e18 cm, 6.e18 cm, 6.e18 cm]
  type: sphere
  number density: 0. cm^-3
  initial temperature: 0. K
"""

WORKLOADS = {
    "lexingtonHII20": dict(grid=64, layout="full", label="benchmarks/lexingtonHII20.param",
                           physics="Planck 20000 K, Verner cross sections, 14 ions + 2 heating terms, Physical diffuse "
                                   "re-emission, temperature solve with line cooling",
                           spinup_packets=1_000_000, packets=100_000_000),
    "stromgren256": dict(grid=256, layout="honly", label="stromgren.param physics on a 256^3 grid (north-star target grid)",
                         physics="monochromatic 13.6 eV, FixedValue cross sections (H only), no diffuse field",
                         spinup_packets=16_000_000, packets=1_000_000_000),
    "clumpy256": dict(grid=256, layout="honly", label="synthetic clumpy 256^3, 16 sources, H-only (BASELINE.json configs[4])",
                      physics="monochromatic 13.6 eV, FixedValue cross sections (H only), no diffuse field",
                      spinup_packets=16_000_000, packets=1_000_000_000),
    "clumpy256L": dict(grid=256, layout="full", label="synthetic clumpy 256^3, 16 sources, Lexington physics",
                       physics="Planck 40000 K, Verner cross sections, 14 ions + 2 heating terms, Physical diffuse "
                               "re-emission, temperature solve with line cooling",
                       spinup_packets=16_000_000, packets=100_000_000),
}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu summaries
    (profiles/traffic.json, written by tools/ncu_traffic.py from a named .ncu-rep): bytes per cell crossing."""
    p = ROOT / "profiles" / "traffic.json"
    if not p.exists():
        return None
    try:
        return json.loads(p.read_text()).get(workload)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int, period: float = 0.2):
        self.device = device
        self.period = float(os.environ.get("BENCH_CLOCK_PERIOD", period))
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.device)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[2 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- reference arm
def cpu_reference_measure(full_packets: int, sample_packets: int, steps: int, warmup: int):
    """Steady-state cost per iteration of the UNMODIFIED reference on this host's cores.

    The reference's IonizationSimulation is built from the Lexington HII20 parameter file and its loop body
    (IonizationSimulation.cpp:359-643) is executed one iteration at a time on the reference's own objects (oracle
    probe cmi_ref_sim_iteration), each phase under its own timer.  `warmup` (>= 5: the temperature solve starts at
    the 5th iteration, TemperatureCalculator.cpp:948) untimed iterations, then `steps` timed ones of
    `sample_packets` packets.  Shoot time is linear in the packet count, reset / re-emission probabilities / state
    update do not depend on it, so the full-size figure is N_full / (N_full / rate_shoot + t_prep + t_update)."""
    import oracle.ref as ref
    ref.use_all_host_threads()  # under torchrun OMP_NUM_THREADS is 1: the reference gets all the host's cores anyway
    warmup = max(warmup, 5)
    with tempfile.TemporaryDirectory() as d:
        yml = Path(d) / "lexingtonHII20.yml"
        yml.write_text(LEXINGTON_YML)
        pf = Path(d) / "lexingtonHII20.param"
        pf.write_text(LEXINGTON_PARAM.format(yml=yml, nc=64, nit=warmup + steps, npk=sample_packets))
        sim = ref.Simulation(pf, num_threads=-1)
        loop = 0
        for _ in range(warmup):
            sim.iteration(loop, min(sample_packets, 200_000))
            loop += 1
        shoot, update, prep = [], [], []
        for _ in range(steps):
            r = sim.iteration(loop, sample_packets)
            loop += 1
            shoot.append(r["shoot_s"]); update.append(r["update_s"]); prep.append(r["prep_s"])
        threads = sim.threads
        sim.close()
    shoot, update, prep = float(np.mean(shoot)), float(np.mean(update)), float(np.mean(prep))
    cpu_model = "unknown"
    try:
        for l in open("/proc/cpuinfo"):
            if l.startswith("model name"):
                cpu_model = l.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    rate_shoot = sample_packets / shoot
    t_full = full_packets / rate_shoot + update + prep
    return {
        "value": full_packets / t_full, "unit": UNIT, "cores": threads, "cpu_model": cpu_model, "kind": "reference",
        "sample": (f"unmodified reference (oracle/_ref, OpenMP, {threads} threads) on lexingtonHII20 64^3: "
                   f"{steps} steady-state iterations of {sample_packets:.0e} packets after {warmup} warm-up "
                   f"iterations (shoot {shoot:.3f} s = {rate_shoot:.3e} packets/s, state update {update:.3f} s, "
                   f"reset+re-emission probabilities {prep:.3f} s per iteration); extrapolated to "
                   f"{full_packets:.0e} packets/iteration: shoot time linear in packets, the rest constant"),
        "shoot_packets_per_s": rate_shoot, "update_s_per_iteration": update, "prep_s_per_iteration": prep,
        "s_per_iteration_at_full_size": t_full,
        # nothing modelled: what the timed iterations themselves did (per-iteration costs weigh 100x more at this size)
        "measured_at_sample_size": {"packets_per_iteration": sample_packets, "s_per_iteration": shoot + update + prep,
                                    "value": sample_packets / (shoot + update + prep), "unit": UNIT},
    }


def _claim_stdout():
    """Libraries (NCCL's version banner, torchrun's OMP notice) write to fd 1; the contract is ONE JSON line on
    stdout.  Point fd 1 at stderr for the duration of the run and keep the real stdout for the result line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def config_of(name, n_packets, world, scaling="strong"):
    w = WORKLOADS[name]
    nc = w["grid"]
    return {"workload": name, "description": w["label"], "grid": f"{nc}^3", "packets_per_iteration": n_packets,
            "physics": w["physics"],
            "parallelism": (f"{world} GPU(s): {scaling} scaling, packets split by global id "
                            "(MPICommunicator::distribute), replicated grid, accumulators all-reduced, "
                            "state update of the owned cell chunks, opacity records all-gathered (NCCL behind the C ABI)"),
            "l2_policy": ("cells + accumulators (41 MB) are L2 resident and re-zeroed / rewritten every iteration; each "
                          "step streams 1e8 independent random rays (up to 21 GB of packet queues per round: larger than L2)"
                          if nc == 64 else
                          "cells + accumulators (0.5 - 2.7 GB) do not fit in L2; every step streams independent random rays")}


# --------------------------------------------------------------------------- our arm
def run_workload(name, n_packets, args, rank, world, local_rank, steps, warmup, spinup, want_e2e, scatter_rates):
    """Time `steps` iterations of workload `name` at `n_packets` packets per iteration (whole job, split over the
    ranks).  Returns the result record of rank 0 (every rank returns the same numbers after the max-reductions)."""
    import torch
    import torch.distributed as dist
    from cmacionize_b200 import capi, problems
    from cmacionize_b200.distributed import cell_block, init_communicator, shard_packets

    w = WORKLOADS[name]
    if name == "lexingtonHII20":
        prob = problems.lexington(20, ncell=w["grid"], n_packets=n_packets, device=local_rank)
    elif name == "stromgren256":
        prob = problems.stromgren(ncell=256, n_packets=n_packets, device=local_rank)
    else:
        prob = problems.synthetic_clumpy(ncell=256, n_packets=n_packets, device=local_rank,
                                         variant="Lexington" if name == "clumpy256L" else "H")
    ctx = prob.ctx
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))
    init_communicator(ctx)                       # no-op for one rank
    lo, cnt = shard_packets(n_packets, rank, world)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = []       # per timed step: events (start, after reset+probabilities, after shoot, after exchange+update)
    kernel_ms = []   # (prepare ms, march ms, rounds) per timed step, from the library's CUDA events
    exch_ms = []     # (reduce, update, gather) of the exchange, N > 1

    def step(loop, npk_total=None, timed=False):
        my_lo, my_cnt = (lo, cnt) if npk_total is None else shard_packets(npk_total, rank, world)
        with torch.cuda.stream(stream):
            e = [ev() for _ in range(4)] if timed else None
            if timed: e[0].record(stream)
            ctx.reset_accumulators()
            ctx.update_reemission_probabilities()
            if timed: e[1].record(stream)
            ctx.shoot(my_cnt, packet_offset=my_lo, seed=prob.seed, iteration=loop, want_counters=False)
            if timed:
                e[2].record(stream)
                kernel_ms.append(ctx.shoot_timing(want_adds=False)[:3] + ctx.shoot_overlap())
            ctx.exchange_and_update(loop)        # N == 1: the state update alone
            if timed:
                e[3].record(stream)
                marks.append(e)
                if world > 1:
                    exch_ms.append(ctx.exchange_timing())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    loop = 0
    for _ in range(spinup):
        step(loop, npk_total=w["spinup_packets"])
        loop += 1
    for _ in range(warmup):
        step(loop)
        loop += 1
    barrier()
    launches0 = capi.kernel_launch_count()
    ctx.set_shoot_timing(True)
    with ClockSampler(local_rank) as clocks:
        t0e, t1e = ev(), ev()
        with torch.cuda.stream(stream):
            t0e.record(stream)
        for _ in range(steps):
            step(loop, timed=True)
            loop += 1
        with torch.cuda.stream(stream):
            t1e.record(stream)
        barrier()
    launches = capi.kernel_launch_count() - launches0
    elapsed = max_over_ranks(t0e.elapsed_time(t1e) * 1e-3)
    ms_per_step = 1e3 * elapsed / steps
    value = n_packets * steps / elapsed
    # counters of the last step: after the exchange every rank holds the sums over the ranks
    crossings, emissions = ctx.shoot_statistics()
    _, _, _, red_ops = ctx.shoot_timing()
    mean = lambda xs: float(np.mean(xs))
    prep_ms, march_ms, rounds, lanes, overlap_ms = (mean([k[i] for k in kernel_ms]) for i in range(5))
    phases = {"reset_and_reemission_probabilities": mean([m[0].elapsed_time(m[1]) for m in marks]),
              "shoot": mean([m[1].elapsed_time(m[2]) for m in marks]),
              "shoot_prepare_kernels": prep_ms, "shoot_march_kernels": march_ms,
              "shoot_prepare_beside_march": overlap_ms,
              "exchange_and_update": mean([m[2].elapsed_time(m[3]) for m in marks])}
    if exch_ms:
        phases["exchange_reduce"], phases["exchange_update_block"], phases["exchange_gather"] = (
            mean([x[i] for x in exch_ms]) for i in range(3))
    phases = {k: max_over_ranks(v) for k, v in phases.items()}
    prep_ms, march_ms = phases["shoot_prepare_kernels"], phases["shoot_march_kernels"]

    # ---- roofline of the dominant kernel (the voxel walk + accumulation) ----
    # unit of work = one packet-cell crossing.  One step launches the kernel once per round of the wavefront
    # pipeline; bytes and time are summed over the rounds of a step (same ratio as per-launch averages).  Times
    # are CUDA events recorded by the library on its own stream.  All counts are per rank.
    peak, peak_src = measured_peaks()
    my_crossings, my_reds = crossings / world, red_ops / world
    red_per_crossing = red_ops / max(crossings, 1.)
    crossing_rate = my_crossings / (march_ms * 1e-3)
    full_layout = w["layout"] == "full"
    gather_bytes = 32. if full_layout else 16.
    survey_bytes = prob.bytes_per_step                      # 152 B (full layout), 24 B (H-only)
    issued_bytes = gather_bytes + 8. * red_per_crossing     # what the kernel really moves per crossing
    l2_resident = w["grid"] <= 64
    red_peak, gather_peak = scatter_rates
    lane_bound = 1. / (1. / gather_peak + red_per_crossing / red_peak)
    hbm_achieved = my_crossings * (issued_bytes if l2_resident else survey_bytes) / (march_ms * 1e-3) / 1e9
    traffic = measured_traffic(name)
    roofline = {
        "kernel": ("march_kernel<ACC_FULL>" if full_layout else
                   ("march_kernel<ACC_HONLY>" if l2_resident else "march_lean_kernel (H-only coherent walk)")),
        "cell_crossings_per_packet": crossings / n_packets, "emissions_per_packet": emissions / n_packets,
        "accumulator_adds_per_crossing": red_per_crossing,
        "algorithmic_bytes_per_crossing": {"survey_8d": survey_bytes, "issued": issued_bytes,
                                           "note": "issued = record gathered + 8 B per accumulator term actually "
                                                   "added (zero terms are skipped, in-warp sums merge same-cell terms); "
                                                   "both counted on the device in this run"},
        "kernel_ms": march_ms, "kernel_launches_per_step": rounds, "kernel_share_of_step": march_ms / ms_per_step,
        "kernel_ms_note": ("time during which a march kernel was running (CUDA events of the library on its streams; "
                           f"{int(round(lanes))} lane(s): with two, the union of the launches' intervals, and the emission "
                           "kernels of the other lane share the SMs during `shoot_prepare_beside_march` ms of it)"),
        "peak_source": peak_src,
        "traffic": None, "traffic_note": None,
    }
    if traffic:
        roofline["traffic"] = traffic["dram_bytes_per_crossing"] * my_crossings / max(rounds, 1.)
        roofline["traffic_note"] = (f"{traffic['dram_bytes_per_crossing']:.2f} DRAM bytes per crossing "
                                    f"(dram__bytes_read.sum + dram__bytes_write.sum of {traffic['kernel']} in "
                                    f"{traffic['report']}, ncu --set full) x the crossings of one launch of this run")
    if l2_resident:
        # the working set (8 MB cells + 34 MB accumulators) is L2 resident: the binding unit is the L1TEX lane
        # throughput of scattered gathers and REDs, measured on this device in this run
        roofline.update({
            "bound": "l1tex-lane", "achieved": crossing_rate, "peak": lane_bound, "unit": "cell crossings/s",
            "frac": crossing_rate / lane_bound,
            "peak_note": (f"1 / (1/gathers_per_s + REDs_per_crossing/REDs_per_s) with {gather_peak:.3e} scattered "
                          f"16-B gathers/s and {red_peak:.3e} scattered FP64 RED/s measured on this device in this run "
                          "(cmib_measure_scatter_rates: every lane of a warp in a different 128-B line)"),
            "hbm": {"achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak,
                    "note": "issued bytes / kernel time against the HBM copy peak: not the binding roof here "
                            "(ncu: DRAM sees the packet queues only)"}})
    else:
        roofline.update({
            "bound": "hbm", "achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak,
            "issued_bytes_frac": my_crossings * issued_bytes / (march_ms * 1e-3) / 1e9 / peak,
            "l1tex_lane": {"achieved": crossing_rate, "peak": lane_bound, "unit": "cell crossings/s",
                           "frac": crossing_rate / lane_bound,
                           "note": "scattered-lane bound of a walk WITHOUT coherence (one gather + the REDs per lane); "
                                   "the coherent walk may exceed it: its lanes share sectors and sum in registers"}})
    rec = {"value": value, "unit": UNIT, "ms_per_step": ms_per_step, "s_per_iteration": ms_per_step * 1e-3,
           "steps": steps, "warmup": warmup, "config": config_of(name, n_packets, world),
           "roofline": roofline, "phases_ms": phases, "gpu_launches": int(launches), "clocks": clocks.summary()}

    # ---- e2e: the same iteration through the C ABI with HOST buffers ----
    # every rank uploads the cells it owns from pinned host memory, their opacity records are gathered over NVLink,
    # the rank shoots its packets, joins the exchange and reads the cells it updated back; wall clock between
    # barriers, max over ranks
    if want_e2e:
        nc = ctx.ncells
        pin = lambda *shape: torch.empty(*shape, dtype=torch.float64).pin_memory().numpy()
        no = ctx.owned_cells()       # the cells this rank updates: chunks dealt round-robin (cmib_owned_cell)
        n_o, T_o, x_o, heat_o = pin(no), pin(no), pin(14, no), pin(2, no)
        if world > 1:
            ctx.comm_gather_state()
        ctx.download_cells_owned_into(n_o, T_o, x_o, heat_o)
        e2e_steps = max(2, min(steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            with torch.cuda.stream(stream):
                ctx.upload_cells_owned(n_o, T_o, x_o)                     # H2D: 16 doubles per owned cell
                ctx.comm_gather_owned_cells()                             # N > 1: replicate the opacity records
                step(loop)
                ctx.download_cells_owned_into(n_o, T_o, x_o, heat_o)      # D2H: 18 doubles per cell the rank updated
            loop += 1
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        rec["e2e"] = {"value": n_packets * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": 16 * 8 * nc,
                      "d2h_bytes_per_step": 18 * 8 * nc, "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
                      "note": "bytes are totals over the ranks: every rank moves the cells it owns"}
    else:
        rec["e2e"] = None
    if world > 1:
        ctx.comm_finalize()
    ctx.close()
    return rec


def main():
    result_out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--packets", type=float, default=FULL_PACKETS, help="packets per iteration of the whole job (benchmark: 1e8)")
    ap.add_argument("--spinup", type=int, default=6, help="untimed iterations with fewer packets that bring the grid to the "
                    "ionised steady state and past the 4 ionization-only iterations")
    ap.add_argument("--cpu-sample", type=float, default=1e6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    ap.add_argument("--workload", default="lexingtonHII20", choices=sorted(WORKLOADS),
                    help="the workload of the top-level keys (default: the configuration BASELINE.json's metric is quoted on)")
    ap.add_argument("--workloads", default="stromgren256,clumpy256",
                    help="comma-separated extra workloads attached under `workloads` ('' = none)")
    ap.add_argument("--extra-steps", type=int, default=3)
    ap.add_argument("--extra-packets", type=float, default=0., help="packets per iteration of the extra workloads "
                    "(default: 1e9, BASELINE.json configs[4])")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_packets = int(args.packets)

    if args.impl == "reference":
        if rank != 0:
            return
        if args.workload != "lexingtonHII20":
            raise SystemExit("the reference arm is timed on the headline workload (lexingtonHII20) only")
        cb = cpu_reference_measure(n_packets, int(args.cpu_sample), args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": max(args.warmup, 5),
                "ms_per_step": 1e3 * cb["s_per_iteration_at_full_size"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_of("lexingtonHII20", n_packets, max(args.gpus, 1)), "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=result_out, flush=True)
        return

    import torch
    import torch.distributed as dist
    from cmacionize_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # scattered-lane ceilings of this device, measured now (denominators of the l1tex-lane roofline)
    with capi.Context([0, 0, 0], [1, 1, 1], [64, 64, 64], device=local_rank) as probe:
        scatter_l2 = probe.measure_scatter_rates(64 ** 3)
        scatter_hbm = probe.measure_scatter_rates(256 ** 3)

    head = run_workload(args.workload, n_packets, args, rank, world, local_rank, args.steps, args.warmup, args.spinup,
                        not args.no_e2e, scatter_l2 if WORKLOADS[args.workload]["grid"] <= 64 else scatter_hbm)
    metric = METRIC if args.workload == "lexingtonHII20" else METRIC.replace("lexingtonHII20", args.workload)
    line = {"metric": metric, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": head["config"],
            "s_per_iteration": head["s_per_iteration"], "roofline": head["roofline"], "phases_ms": head["phases_ms"],
            "gpu_launches": head["gpu_launches"], "clocks": head["clocks"], "e2e": head["e2e"],
            "scatter_rates_measured": {"l2_resident_table_64^3": {"red_per_s": scatter_l2[0], "gather_per_s": scatter_l2[1]},
                                       "hbm_resident_table_256^3": {"red_per_s": scatter_hbm[0], "gather_per_s": scatter_hbm[1]}}}

    if world > 1 and not args.no_weak:
        # the weak figure (every GPU shoots the configuration's packet count: N x 1e8 distinct packets per iteration)
        wk = run_workload(args.workload, n_packets * world, args, rank, world, local_rank, 2, 1, args.spinup, False,
                          scatter_l2 if WORKLOADS[args.workload]["grid"] <= 64 else scatter_hbm)
        line["weak"] = {"value": wk["value"], "unit": UNIT, "ms_per_step": wk["ms_per_step"],
                        "packets_per_iteration": n_packets * world, "steps": 2, "phases_ms": wk["phases_ms"]}

    extra = [x for x in args.workloads.split(",") if x and x != args.workload]
    if extra:
        line["workloads"] = {}
        for name in extra:
            rec = run_workload(name, int(args.extra_packets or WORKLOADS[name]["packets"]), args, rank, world, local_rank,
                               args.extra_steps, 2, args.spinup,
                               not args.no_e2e, scatter_l2 if WORKLOADS[name]["grid"] <= 64 else scatter_hbm)
            rec["metric"] = METRIC.replace("lexingtonHII20", name)
            rec["n_gpus"] = world
            rec["scaling"] = "strong"
            line["workloads"][name] = rec

    if rank == 0 and not args.no_cpu_baseline and world == 1 and args.workload == "lexingtonHII20":
        try:
            line["cpu_baseline"] = cpu_reference_measure(n_packets, int(args.cpu_sample), 2, 5)
        except Exception as exc:  # the oracle is optional evidence, never part of the product path
            line["cpu_baseline"] = {"error": repr(exc)}

    if rank == 0:
        print(json.dumps(line), file=result_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
