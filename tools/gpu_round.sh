#!/bin/bash
# One GPU box visit: parity tests, bench, launch list, full ncu capture of the hot kernels.
# Usage (under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/gpu_tests_$TAG.log
tail -8 $OUT/gpu_tests_$TAG.log | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cat $OUT/bench_$TAG.json | cut -c1-1200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --spinup 5 --packets 1e7 --no-cpu-baseline --no-e2e \
  > $OUT/bench_under_ncu_$TAG.log 2>&1
echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:march_kernel|prepare_kernel' --launch-skip 340 --launch-count 12 \
  -f -o $OUT/wavefront_full_$TAG python bench.py --steps 1 --warmup 3 --spinup 5 --packets 1e7 --no-cpu-baseline --no-e2e \
  > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT
