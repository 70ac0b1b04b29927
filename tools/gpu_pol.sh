#!/bin/bash
# r02: A/B of one build of the coherent walk on the 256^3 grids (compare with the previous build's lines in profiles/)
TAG=${1:-pol}
OUT=gpurun_out
mkdir -p $OUT
run() { # problem packets repeat env...
  local prob=$1 n=$2 rep=$3; shift 3
  echo "## $prob $n $*" >> $OUT/ab_$TAG.txt
  env "$@" timeout 300 python tools/profile_shoot.py --problem $prob --packets $n --repeat $rep --spinup-packets 16000000 2>&1 | grep -v "^$" | tail -$((rep-1)) | cut -c1-380 >> $OUT/ab_$TAG.txt
}
: > $OUT/ab_$TAG.txt
for prob in stromgren256 clumpy256; do
  run $prob 125000000 3 CMIB_X=0
  run $prob 16000000 3 CMIB_X=0
done
cat $OUT/ab_$TAG.txt
