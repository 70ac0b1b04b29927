#!/bin/bash
# Full GPU check of the current tree + ncu of the coherent march on the HBM-resident synthetic grid.
TAG=${1:-r01q}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -4 $OUT/gpu_tests_$TAG.log | cut -c1-300
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cut -c1-300 $OUT/bench_$TAG.json
for prob in clumpy256 clumpy256L; do
  CMIB_SORT=2 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k 'regex:^march_kernel' --launch-count 1 -f -o $OUT/coherent_${prob}_$TAG python tools/profile_shoot.py --problem $prob --packets 16000000 > $OUT/ncu_coherent_${prob}_$TAG.log 2>&1
  echo "ncu $prob rc=$?"
done
CMIB_SORT=0 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k 'regex:^march_kernel' --launch-count 1 -f -o $OUT/plain_clumpy256_$TAG python tools/profile_shoot.py --problem clumpy256 --packets 16000000 > $OUT/ncu_plain_clumpy256_$TAG.log 2>&1
echo "ncu plain rc=$?"
