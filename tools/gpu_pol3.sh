#!/bin/bash
TAG=${1:-pol3}
OUT=gpurun_out
run() { local prob=$1 n=$2 rep=$3; shift 3
  echo "## $prob $n $*" >> $OUT/ab_$TAG.txt
  env "$@" timeout 300 python tools/profile_shoot.py --problem $prob --packets $n --repeat $rep --spinup-packets ${SPIN:-16000000} 2>&1 | grep -v "^$" | tail -$((rep-1)) | cut -c1-380 >> $OUT/ab_$TAG.txt
}
: > $OUT/ab_$TAG.txt
run stromgren256 125000000 3 CMIB_X=0
run stromgren256 16000000 3 CMIB_X=0
run clumpy256 125000000 3 CMIB_X=0
SPIN=2000000 run lexington 100000000 3 CMIB_X=0
cat $OUT/ab_$TAG.txt
