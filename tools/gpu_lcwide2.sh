#!/bin/bash
TAG=${1:-lcw2}
OUT=gpurun_out
for w in 0 3 9 0 3 9; do
  CMIB_LC_WIDE=$w timeout 600 python bench.py --steps 24 --warmup 3 --no-e2e --no-cpu-baseline --workloads '' > $OUT/bench_lcwide${w}_$TAG.json 2> $OUT/bench_lcwide${w}_$TAG.err
  python - <<P
import json
d = json.loads(open("$OUT/bench_lcwide${w}_$TAG.json").read().strip().splitlines()[-1])
print("CMIB_LC_WIDE=$w", "%.2f ms/step" % d["ms_per_step"], "update %.3f" % d["phases_ms"]["exchange_and_update"], "shoot %.2f" % d["phases_ms"]["shoot"])
P
done
