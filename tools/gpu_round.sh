#!/bin/bash
# One GPU box visit: parity tests, bench, launch list, full ncu capture of the shoot kernel.
# Usage (under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/gpu_tests_$TAG.log
tail -5 $OUT/gpu_tests_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cat $OUT/bench_$TAG.json | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --packets 1e7 --no-cpu-baseline --no-e2e \
  > $OUT/bench_under_ncu_$TAG.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:shoot_kernel --launch-skip 9 --launch-count 1 \
  -f -o $OUT/shoot_full_$TAG python bench.py --steps 2 --warmup 3 --packets 1e7 --no-cpu-baseline --no-e2e \
  > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT
