#!/bin/bash
# What the driver does at round end, in one visit: build + smoke, GPU tests, both bench arms.
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke_$TAG.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/gpu_tests_$TAG.log; tail -4 $OUT/gpu_tests_$TAG.log | cut -c1-300
cp $OUT/parity_physics.json $OUT/parity_physics_$TAG.json 2>/dev/null; cp $OUT/parity_lexington.json $OUT/parity_lexington_$TAG.json 2>/dev/null
timeout 900 python bench.py --impl reference --steps 2 --warmup 5 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-300 $OUT/bench_ref_$TAG.json
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cut -c1-400 $OUT/bench_$TAG.json; tail -2 $OUT/bench_$TAG.err
