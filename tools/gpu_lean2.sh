#!/bin/bash
# r02: counting sort + optical-depth bins in the key + march_lean_kernel v3; A/B on the 256^3 grids and 64^3 stromgren
TAG=${1:-lean2}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_shoot.py tests/test_gpu_march.py tests/test_gpu_simulation.py -m gpu -q -x --timeout 600 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/gpu_tests_$TAG.log | cut -c1-300
run() { # problem packets repeat env...
  local prob=$1 n=$2 rep=$3; shift 3
  echo "## $prob $n $*" >> $OUT/ab_$TAG.txt
  env "$@" timeout 300 python tools/profile_shoot.py --problem $prob --packets $n --repeat $rep --spinup-packets 16000000 2>&1 | grep -v "^$" | tail -$((rep-1)) | cut -c1-330 >> $OUT/ab_$TAG.txt
}
: > $OUT/ab_$TAG.txt
for prob in "stromgren256 16000000" "clumpy256 16000000"; do
  set -- $prob
  run $1 $2 3 CMIB_SORT=2 CMIB_TAU_BITS=0
  run $1 $2 3 CMIB_SORT=2 CMIB_TAU_BITS=2
  run $1 $2 3 CMIB_SORT=2 CMIB_TAU_BITS=4
  run $1 $2 3 CMIB_SORT=2 CMIB_TAU_BITS=6
  run $1 $2 3 CMIB_SORT=2 CMIB_TAU_BITS=4 CMIB_LEAN_STEPS=2
  run $1 $2 3 CMIB_SORT=2 CMIB_TAU_BITS=4 CMIB_PREFETCH=0
done
run clumpy256L 16000000 2 CMIB_SORT=2 CMIB_TAU_BITS=0
run clumpy256L 16000000 2 CMIB_SORT=2 CMIB_TAU_BITS=4
run stromgren 4000000 3 CMIB_SORT=0
run stromgren 4000000 3 CMIB_SORT=2 CMIB_TAU_BITS=4
run stromgren 4000000 3 CMIB_SORT=2 CMIB_TAU_BITS=0
cat $OUT/ab_$TAG.txt
