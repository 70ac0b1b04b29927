#!/bin/bash
# r02: warp-wide line cooling for the last cells of a warp (update_temperature_kernel) against the per-lane path
TAG=${1:-lcwide}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_simulation.py tests/test_gpu_physics.py -m gpu -q -x --timeout 900 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/gpu_tests_$TAG.log | cut -c1-300
for w in 0 9 18 30; do
  CMIB_LC_WIDE=$w timeout 600 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --workloads '' > $OUT/bench_lcwide${w}_$TAG.json 2> $OUT/bench_lcwide${w}_$TAG.err
  python - <<P
import json
d = json.loads(open("$OUT/bench_lcwide${w}_$TAG.json").read().strip().splitlines()[-1])
print("CMIB_LC_WIDE=$w", "%.2f ms/step" % d["ms_per_step"], {k: round(v, 3) for k, v in d["phases_ms"].items()})
P
done
