#!/bin/bash
# the 256^3 workloads (north-star target grids) on one GPU, final build
TAG=${1:-w256}
OUT=gpurun_out
mkdir -p $OUT
for wl in stromgren256 clumpy256; do
  timeout 500 python bench.py --workload $wl --steps 3 --warmup 3 > $OUT/bench_${wl}_$TAG.json 2> $OUT/bench_${wl}_$TAG.err
  echo "$wl rc=$?"; python - <<P
import json
d=json.loads(open("$OUT/bench_${wl}_$TAG.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("%s: %.1f ms/step %.3e packets/s e2e %s march %.1f ms l1tex %.2f hbm %.2f" % ("$wl", d["ms_per_step"], d["value"], d["e2e"] and "%.3e" % d["e2e"]["value"], r["kernel_ms"], r["l1tex"]["frac"], r["frac"]))
P
done
