#!/bin/bash
# ncu --set full of the H-only coherent walk (march_lean_kernel) on the 256^3 grids
TAG=${1:-lean}
OUT=gpurun_out
mkdir -p $OUT
for prob in stromgren256 clumpy256; do
  CMIB_SORT=2 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:march_lean_kernel' --launch-count 1 \
    -f -o $OUT/lean_${prob}_$TAG python tools/profile_shoot.py --problem $prob --packets 16000000 --spinup-packets 16000000 > $OUT/ncu_lean_${prob}_$TAG.log 2>&1
  echo "ncu $prob rc=$?"; ls -la $OUT/lean_${prob}_$TAG.ncu-rep; tail -2 $OUT/ncu_lean_${prob}_$TAG.log | cut -c1-300
done
