#!/bin/bash
# r02: key layout rule (direction bins from the measured walk length, rest = optical-depth bins); queue capacity; ncu
TAG=${1:-lean3}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_shoot.py tests/test_gpu_march.py -m gpu -q -x --timeout 600 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/gpu_tests_$TAG.log | cut -c1-300
run() { # problem packets repeat env...
  local prob=$1 n=$2 rep=$3; shift 3
  echo "## $prob $n $*" >> $OUT/ab_$TAG.txt
  env "$@" timeout 300 python tools/profile_shoot.py --problem $prob --packets $n --repeat $rep --spinup-packets 16000000 2>&1 | grep -v "^$" | tail -$((rep-1)) | cut -c1-330 >> $OUT/ab_$TAG.txt
}
: > $OUT/ab_$TAG.txt
for prob in "stromgren256" "clumpy256"; do
  run $prob 16000000 3 CMIB_SORT=2
  run $prob 16000000 3 CMIB_SORT=2 CMIB_TAU_BITS=8 CMIB_DIR_BITS=14
  run $prob 16000000 3 CMIB_SORT=2 CMIB_KEY_BITS=24
  run $prob 64000000 3 CMIB_SORT=2 CMIB_QUEUE_CAPACITY=67108864
  run $prob 64000000 3 CMIB_SORT=2
done
run clumpy256L 16000000 2 CMIB_SORT=2
CMIB_SORT=2 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:march_lean_kernel' --launch-count 1 \
    -f -o $OUT/lean_stromgren256_$TAG python tools/profile_shoot.py --problem stromgren256 --packets 16000000 --spinup-packets 16000000 > $OUT/ncu_lean_stromgren256_$TAG.log 2>&1
echo "ncu rc=$?"
cat $OUT/ab_$TAG.txt
