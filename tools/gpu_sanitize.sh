#!/bin/bash
# compute-sanitizer over the small end-to-end target, both queue orders
TAG=${1:-san}
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck initcheck racecheck synccheck; do
  for sort in 0 2; do
    CMIB_SORT=$sort timeout 100 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > $OUT/sanitize_${tool}_sort${sort}_$TAG.log 2>&1
    echo "$tool sort=$sort rc=$? $(grep -E 'ERROR SUMMARY|sanitize target done' $OUT/sanitize_${tool}_sort${sort}_$TAG.log | tr '\n' ' ')"
  done
done
