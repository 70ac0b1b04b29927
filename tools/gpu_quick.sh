#!/bin/bash
# quick GPU visit: gpu tests + profile_shoot timings + bench
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/gpu_tests_$TAG.log
tail -12 $OUT/gpu_tests_$TAG.log | cut -c1-300
python tools/profile_shoot.py --repeat 3 2>&1 | tail -2 | tee $OUT/profile_shoot_$TAG.txt
python tools/profile_shoot.py --repeat 3 --problem stromgren --packets 4e6 2>&1 | tail -2 | tee -a $OUT/profile_shoot_$TAG.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cut -c1-700 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
