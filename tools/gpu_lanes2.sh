#!/bin/bash
# r02: lanes at the packet counts of the 256^3 workloads (1e9 per iteration: 1 GPU; 1.25e8: one of 8 GPUs)
TAG=${1:-lanes2}
OUT=gpurun_out
mkdir -p $OUT
run() { # problem packets repeat env...
  local prob=$1 n=$2 rep=$3; shift 3
  echo "## $prob $n $*" >> $OUT/ab_$TAG.txt
  env "$@" timeout 300 python tools/profile_shoot.py --problem $prob --packets $n --repeat $rep --spinup-packets 16000000 2>&1 | grep -v "^$" | tail -$((rep-1)) | cut -c1-380 >> $OUT/ab_$TAG.txt
}
: > $OUT/ab_$TAG.txt
for prob in stromgren256 clumpy256; do
  for n in 1000000000 125000000; do
    run $prob $n 3 CMIB_LANES=1
    run $prob $n 3 CMIB_LANES=2
    run $prob $n 3 CMIB_LANES=2 CMIB_QUEUE_CAPACITY=33554432
  done
done
run stromgren256 1000000000 3 CMIB_LANES=2 CMIB_QUEUE_CAPACITY=16777216
run stromgren256 125000000 3 CMIB_LANES=2 CMIB_QUEUE_CAPACITY=16777216
cat $OUT/ab_$TAG.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_1gpu_$TAG.json 2> $OUT/bench_1gpu_$TAG.err
echo "bench1 rc=$?"; tail -3 $OUT/bench_1gpu_$TAG.err | cut -c1-300
