#!/bin/bash
# r02: does the ordered queue pay on the L2-resident lexingtonHII20 grid with the r02 key (direction + depth bits)?
TAG=${1:-lexsort}
OUT=gpurun_out
mkdir -p $OUT
run() { # problem packets repeat env...
  local prob=$1 n=$2 rep=$3; shift 3
  echo "## $prob $n $*" >> $OUT/ab_$TAG.txt
  env "$@" timeout 300 python tools/profile_shoot.py --problem $prob --packets $n --repeat $rep --spinup-packets 2000000 2>&1 | grep -v "^$" | tail -$((rep-1)) | cut -c1-380 >> $OUT/ab_$TAG.txt
}
: > $OUT/ab_$TAG.txt
run lexington 100000000 3 CMIB_SORT=0
run lexington 100000000 3 CMIB_SORT=1
run lexington 100000000 3 CMIB_SORT=2
run lexington 100000000 3 CMIB_SORT=1 CMIB_SORT_REEMITTED=1
run lexington 100000000 3 CMIB_SORT=2 CMIB_SORT_REEMITTED=1
run lexington 100000000 3 CMIB_SORT=1 CMIB_DIR_BITS=12
run lexington 100000000 3 CMIB_SORT=2 CMIB_DIR_BITS=12
run lexington 100000000 3 CMIB_SORT=2 CMIB_DIR_BITS=14 CMIB_TAU_BITS=8
run stromgren 100000000 3 CMIB_SORT=0
run stromgren 100000000 3 CMIB_SORT=2
cat $OUT/ab_$TAG.txt
