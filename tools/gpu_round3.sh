#!/bin/bash
# Round-end check (build + smoke, GPU tests, both bench arms) + the 256^3 workloads on one GPU.
TAG=${1:-r01s}
OUT=gpurun_out
mkdir -p $OUT
bash tools/gpu_final.sh $TAG
for wl in stromgren256 clumpy256 clumpy256L; do
  timeout 900 python bench.py --workload $wl --no-e2e > $OUT/bench_${wl}_$TAG.json 2> $OUT/bench_${wl}_$TAG.err
  echo "bench $wl rc=$?"; cut -c1-330 $OUT/bench_${wl}_$TAG.json
done
CMIB_SORT=2 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k 'regex:^march_kernel' --launch-count 1 -f -o $OUT/coherent_clumpy256_$TAG python tools/profile_shoot.py --problem clumpy256 --packets 16000000 > $OUT/ncu_coherent_clumpy256_$TAG.log 2>&1
echo "ncu rc=$?"
