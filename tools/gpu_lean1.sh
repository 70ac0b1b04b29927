#!/bin/bash
# r02: A/B of the H-only coherent walk (march_lean_kernel, march_coherent.cuh) against the r01 kernels
TAG=${1:-lean1}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_shoot.py tests/test_gpu_march.py -m gpu -q -x --timeout 600 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/gpu_tests_$TAG.log | cut -c1-300
run() { # problem packets repeat env...
  local prob=$1 n=$2 rep=$3; shift 3
  echo "## $prob $n $*" >> $OUT/ab_$TAG.txt
  env "$@" timeout 300 python tools/profile_shoot.py --problem $prob --packets $n --repeat $rep --spinup-packets 16000000 2>&1 | grep -v "^$" | tail -$((rep-1)) | cut -c1-330 >> $OUT/ab_$TAG.txt
}
: > $OUT/ab_$TAG.txt
for prob in "stromgren256 16000000" "clumpy256 16000000"; do
  set -- $prob
  run $1 $2 3 CMIB_SORT=0
  run $1 $2 3 CMIB_SORT=2 CMIB_LEAN=0
  run $1 $2 3 CMIB_SORT=2 CMIB_LEAN=1
  run $1 $2 3 CMIB_SORT=2 CMIB_LEAN=1 CMIB_PREFETCH=0
  run $1 $2 3 CMIB_SORT=2 CMIB_LEAN=1 CMIB_LEAN_STEPS=2
  run $1 $2 3 CMIB_SORT=2 CMIB_LEAN=1 CMIB_MARCH_BLOCKS_PER_SM=3
done
cat $OUT/ab_$TAG.txt
