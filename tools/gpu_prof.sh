#!/bin/bash
TAG=${1:-p}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:^(march|prepare)_kernel' --launch-count 4 \
  -f -o $OUT/wavefront_lex_$TAG python tools/profile_shoot.py > $OUT/ncu_lex_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/ncu_lex_$TAG.log
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:^(march|prepare)_kernel' --launch-count 2 \
  -f -o $OUT/wavefront_strom_$TAG python tools/profile_shoot.py --problem stromgren --packets 4e6 > $OUT/ncu_strom_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/ncu_strom_$TAG.log
