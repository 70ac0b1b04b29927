#!/bin/bash
# A/B of coherent-march variants: CTAs per SM the kernel is compiled for (libcmib_agg2.so = 2), prefetch
TAG=${1:-agg}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_shoot.py tests/test_gpu_march.py -m gpu -q -x --timeout 600 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/gpu_tests_$TAG.log | cut -c1-300
run() { # problem packets repeat env...
  local prob=$1 n=$2 rep=$3; shift 3
  echo "## $prob $n $*" >> $OUT/agg_$TAG.txt
  env "$@" timeout 300 python tools/profile_shoot.py --problem $prob --packets $n --repeat $rep 2>&1 | grep -v "^$" | tail -$((rep-1)) | cut -c1-330 >> $OUT/agg_$TAG.txt
}
: > $OUT/agg_$TAG.txt
L2=$PWD/cmacionize_b200/libcmib_agg2.so
for prob in "stromgren256 16000000" "clumpy256 16000000" "clumpy256L 16000000" "lexington 16777216"; do
  set -- $prob
  run $1 $2 2 CMIB_SORT=0
  run $1 $2 2 CMIB_SORT=2
  run $1 $2 2 CMIB_SORT=2 CMIB_PREFETCH=0
  run $1 $2 2 CMIB_SORT=2 CMIB_LIB=$L2
  run $1 $2 2 CMIB_SORT=2 CMIB_LIB=$L2 CMIB_PREFETCH=0
done
cat $OUT/agg_$TAG.txt
