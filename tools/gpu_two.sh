#!/bin/bash
# 2-GPU visit: the 2-GPU host-driver test, then both bench arms as the driver launches them at N=2
TAG=${1:-two}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_host_driver.py -m gpu -q --timeout 500 > $OUT/gpu_tests_2gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -2 $OUT/gpu_tests_2gpu_$TAG.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/bench_2gpu_$TAG.json 2> $OUT/bench_2gpu_$TAG.err
echo "bench rc=$?"; cut -c1-300 $OUT/bench_2gpu_$TAG.json; grep -o '"e2e": {[^}]*}' $OUT/bench_2gpu_$TAG.json; tail -3 $OUT/bench_2gpu_$TAG.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus 2 --steps 2 --warmup 5 > $OUT/bench_ref_2gpu_$TAG.json 2> $OUT/bench_ref_2gpu_$TAG.err
echo "ref rc=$?"; grep -o '"value": [0-9.e+]*\|"cores": [0-9]*' $OUT/bench_ref_2gpu_$TAG.json | head -3
