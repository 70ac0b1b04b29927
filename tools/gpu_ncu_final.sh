#!/bin/bash
# one ncu --set full capture of the wavefront kernels of the final build (first rounds of a 16 Mi packet shoot)
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
timeout 260 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:^(march|prepare|reemit_decide)_kernel' --launch-count 5 \
  -f -o $OUT/wavefront_lex_$TAG python tools/profile_shoot.py --packets 16777216 > $OUT/ncu_lex_$TAG.log 2>&1
echo "ncu rc=$?"; ls -la $OUT/wavefront_lex_$TAG.ncu-rep; tail -2 $OUT/ncu_lex_$TAG.log
