#!/bin/bash
# 8-GPU weak-scaling lines: headline workload + the 256^3 north-star grids (1e8 packets per GPU per iteration)
TAG=${1:-r01v}
OUT=gpurun_out
mkdir -p $OUT
N=${2:-8}
for wl in lexingtonHII20 stromgren256 clumpy256; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload $wl --no-e2e --no-cpu-baseline > $OUT/bench_${N}gpu_${wl}_$TAG.json 2> $OUT/bench_${N}gpu_${wl}_$TAG.err
  echo "bench $N x $wl rc=$?"; cut -c1-220 $OUT/bench_${N}gpu_${wl}_$TAG.json
done
