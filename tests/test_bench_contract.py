"""bench.py's reference arm (the one leg that runs without a GPU): ONE JSON line with the contract's keys, the
reference on all host cores — also under torch.distributed.run, which exports OMP_NUM_THREADS=1 to its workers and
where rank 0 alone may run and print."""
import json
import os
import subprocess
import sys

from conftest import ROOT

KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _check(line, n_gpus):
    d = json.loads(line)
    assert KEYS <= set(d) and d["impl"] == "reference" and d["n_gpus"] == n_gpus
    assert d["unit"] == "packets/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    # strong scaling, as in the reference: the iteration keeps the parameter file's 1e8 packets for every rank count
    assert d["config"]["workload"] == "lexingtonHII20" and d["config"]["packets_per_iteration"] == 100_000_000
    assert d["scaling"] == "strong"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == d["value"] and cb["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": "packets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e4 and abs(d["ms_per_step"] * 1e-3 * d["value"] / d["config"]["packets_per_iteration"] - 1.) < 1e-9


def test_reference_arm_prints_one_contract_line(ref):
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-sample", "2e5"], cwd=str(ROOT), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    _check(lines[0], 1)


def test_reference_arm_under_torchrun_uses_all_cores_and_prints_once(ref):
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", "bench.py", "--impl", "reference",
                          "--gpus", "2", "--steps", "1", "--warmup", "1", "--cpu-sample", "2e5"],
                         cwd=str(ROOT), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    _check(lines[0], 2)
