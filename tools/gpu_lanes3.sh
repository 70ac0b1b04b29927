#!/bin/bash
# r02: the lane / round-size rule as built: GPU tier of the shoot, the 256^3 workloads by default, bench at N = 1
TAG=${1:-lanes3}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_shoot.py tests/test_gpu_march.py tests/test_gpu_parity256.py tests/test_gpu_simulation.py -m gpu -q -x --timeout 900 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/gpu_tests_$TAG.log | cut -c1-300
run() { # problem packets repeat env...
  local prob=$1 n=$2 rep=$3; shift 3
  echo "## $prob $n $*" >> $OUT/ab_$TAG.txt
  env "$@" timeout 300 python tools/profile_shoot.py --problem $prob --packets $n --repeat $rep --spinup-packets 16000000 2>&1 | grep -v "^$" | tail -$((rep-1)) | cut -c1-380 >> $OUT/ab_$TAG.txt
}
: > $OUT/ab_$TAG.txt
for prob in stromgren256 clumpy256; do
  for n in 1000000000 125000000 12500000; do
    run $prob $n 3 CMIB_X=0
  done
done
cat $OUT/ab_$TAG.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_1gpu_$TAG.json 2> $OUT/bench_1gpu_$TAG.err
echo "bench1 rc=$?"; tail -3 $OUT/bench_1gpu_$TAG.err | cut -c1-300
