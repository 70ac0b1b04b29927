#!/bin/bash
# A/B of march-kernel variants (queue order CMIB_SORT, next-cell prefetch CMIB_PREFETCH): correctness first, then timings.
TAG=${1:-agg}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_shoot.py tests/test_gpu_march.py tests/test_gpu_continuous.py -m gpu -q -x --timeout 900 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -5 $OUT/gpu_tests_$TAG.log | cut -c1-400
run() { # problem packets repeat env...
  local prob=$1 n=$2 rep=$3; shift 3
  echo "## $prob $n $*" >> $OUT/agg_$TAG.txt
  env "$@" timeout 300 python tools/profile_shoot.py --problem $prob --packets $n --repeat $rep 2>&1 | grep -v "^$" | tail -$((rep-1)) | cut -c1-330 >> $OUT/agg_$TAG.txt
}
: > $OUT/agg_$TAG.txt
for prob in "stromgren 4000000" "lexington 16777216" "stromgren256 16000000" "clumpy256 16000000" "clumpy256L 16000000"; do
  set -- $prob
  for pre in 0 1; do
    run $1 $2 2 CMIB_SORT=0 CMIB_PREFETCH=$pre
    run $1 $2 2 CMIB_SORT=2 CMIB_PREFETCH=$pre
  done
done
cat $OUT/agg_$TAG.txt
