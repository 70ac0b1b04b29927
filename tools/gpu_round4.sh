#!/bin/bash
TAG=${1:-r01t}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -4 $OUT/gpu_tests_$TAG.log | cut -c1-300
for wl in stromgren256 clumpy256; do
  timeout 900 python bench.py --workload $wl --no-e2e > $OUT/bench_${wl}_$TAG.json 2> $OUT/bench_${wl}_$TAG.err
  echo "bench $wl rc=$?"; cut -c1-200 $OUT/bench_${wl}_$TAG.json
done
