#!/bin/bash
TAG=${1:-p}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --spinup 5 --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu_$TAG.log 2>&1
echo "list rc=$?"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:^(march|prepare|reemit_decide)_kernel' --launch-count 5 \
  -f -o $OUT/wavefront_lex_$TAG python tools/profile_shoot.py --packets 16777216 > $OUT/ncu_lex_$TAG.log 2>&1
echo "ncu rc=$?"; ls -la $OUT | tail -5
