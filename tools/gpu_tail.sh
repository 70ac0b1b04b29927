#!/bin/bash
# r02: tail kernel (last generations of re-emitted packets in one launch) against the round-by-round tail
TAG=${1:-tail}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_shoot.py tests/test_gpu_simulation.py tests/test_gpu_continuous.py -m gpu -q -x --timeout 900 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -5 $OUT/gpu_tests_$TAG.log | cut -c1-300
run() { # problem packets repeat env...
  local prob=$1 n=$2 rep=$3; shift 3
  echo "## $prob $n $*" >> $OUT/ab_$TAG.txt
  env "$@" timeout 300 python tools/profile_shoot.py --problem $prob --packets $n --repeat $rep --spinup-packets 2000000 2>&1 | grep -v "^$" | tail -$((rep-1)) | cut -c1-380 >> $OUT/ab_$TAG.txt
}
: > $OUT/ab_$TAG.txt
for t in 0 1; do
  run lexington 100000000 3 CMIB_TAIL=$t
  run lexington 12500000 4 CMIB_TAIL=$t
done
cat $OUT/ab_$TAG.txt
