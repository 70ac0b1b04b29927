#!/bin/bash
TAG=${1:-p}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_physics.py tests/test_gpu_simulation.py -m gpu -q --timeout 600 2>&1 | tail -6 | cut -c1-300
cp $OUT/parity_physics.json $OUT/parity_physics_$TAG.json
python tools/profile_shoot.py --repeat 2 2>&1 | tail -1
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cut -c1-300 $OUT/bench_$TAG.json
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:^update_state_kernel' --launch-skip 6 --launch-count 1 \
  -f -o $OUT/update_$TAG python tools/profile_shoot.py --packets 1e6 > $OUT/ncu_update_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/ncu_update_$TAG.log
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:^(march|prepare)_kernel' --launch-count 4 \
  -f -o $OUT/wavefront_lex_$TAG python tools/profile_shoot.py > $OUT/ncu_lex_$TAG.log 2>&1
echo "ncu rc=$?"
