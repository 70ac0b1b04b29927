#!/bin/bash
# compute-sanitizer over the small end-to-end target: both queue orders, and two lanes on the ordered queue
TAG=${1:-san}
OUT=gpurun_out
mkdir -p $OUT
run() { # tool name env...
  local tool=$1 name=$2; shift 2
  env "$@" timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > $OUT/sanitize_${tool}_${name}_$TAG.log 2>&1
  echo "$tool $name rc=$? $(grep -E 'ERROR SUMMARY|sanitize target done' $OUT/sanitize_${tool}_${name}_$TAG.log | tr '\n' ' ')"
}
run memcheck sort0 CMIB_SORT=0
run memcheck sort2 CMIB_SORT=2
run memcheck sort2_lanes2 CMIB_SORT=2 CMIB_LANES=2
run racecheck sort2_lanes2 CMIB_SORT=2 CMIB_LANES=2
run racecheck sort0 CMIB_SORT=0
run synccheck sort2_lanes2 CMIB_SORT=2 CMIB_LANES=2
run initcheck sort2_lanes2 CMIB_SORT=2 CMIB_LANES=2
