#!/bin/bash
# r02: two lanes (emission of one beside the walk of the other) against one lane
TAG=${1:-lanes1}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_shoot.py tests/test_gpu_march.py tests/test_gpu_simulation.py -m gpu -q -x --timeout 600 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/gpu_tests_$TAG.log | cut -c1-300
run() { # problem packets repeat env...
  local prob=$1 n=$2 rep=$3; shift 3
  echo "## $prob $n $*" >> $OUT/ab_$TAG.txt
  env "$@" timeout 300 python tools/profile_shoot.py --problem $prob --packets $n --repeat $rep --spinup-packets 16000000 2>&1 | grep -v "^$" | tail -$((rep-1)) | cut -c1-380 >> $OUT/ab_$TAG.txt
}
: > $OUT/ab_$TAG.txt
for lanes in 1 2; do
  run lexington 100000000 3 CMIB_LANES=$lanes
  run stromgren256 100000000 3 CMIB_LANES=$lanes
  run clumpy256 100000000 3 CMIB_LANES=$lanes
  run stromgren256 12500000 3 CMIB_LANES=$lanes
done
run stromgren256 100000000 3 CMIB_LANES=2 CMIB_LANE_ROUNDS=4
run stromgren256 100000000 3 CMIB_LANES=2 CMIB_QUEUE_CAPACITY=16777216
run stromgren256 12500000 3 CMIB_LANES=2 CMIB_LANE_ROUNDS=2
run stromgren256 12500000 3 CMIB_LANES=2 CMIB_LANE_ROUNDS=4
run lexington 100000000 3 CMIB_LANES=2 CMIB_QUEUE_CAPACITY=8388608
cat $OUT/ab_$TAG.txt
