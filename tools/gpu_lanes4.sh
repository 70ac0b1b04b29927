#!/bin/bash
TAG=${1:-lanes4}
OUT=gpurun_out
run() { local prob=$1 n=$2 rep=$3; shift 3
  echo "## $prob $n $*" >> $OUT/ab_$TAG.txt
  env "$@" timeout 300 python tools/profile_shoot.py --problem $prob --packets $n --repeat $rep --spinup-packets 16000000 2>&1 | grep -v "^$" | tail -$((rep-1)) | cut -c1-380 >> $OUT/ab_$TAG.txt
}
: > $OUT/ab_$TAG.txt
for n in 1000000000 125000000; do
  run stromgren256 $n 3 CMIB_LANES=1
  run stromgren256 $n 3 CMIB_LANES=2
done
run stromgren256 1000000000 3 CMIB_LANES=1 CMIB_QUEUE_CAPACITY=33554432
cat $OUT/ab_$TAG.txt
