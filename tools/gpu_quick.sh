#!/bin/bash
# quick GPU visit: microbench + gpu tests (no -x) + short bench
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
[ -x tools/microbench/red_bench ] && timeout 300 tools/microbench/red_bench > $OUT/red_bench_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/gpu_tests_$TAG.log
tail -15 $OUT/gpu_tests_$TAG.log | cut -c1-400
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cut -c1-900 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
