#!/bin/bash
# Round-end visit: what the driver does (build + smoke, GPU tests, both bench arms) + the ncu launch list of the bench.
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
bash tools/gpu_final.sh $TAG
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/launches_bench_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --spinup 5 --no-cpu-baseline --no-e2e > $OUT/ncu_bench_$TAG.log 2>&1
echo "ncu rc=$?"; wc -l $OUT/launches_bench_$TAG.csv
