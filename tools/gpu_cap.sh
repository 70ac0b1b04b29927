#!/bin/bash
OUT=gpurun_out; TAG=${1:-cap1}
: > $OUT/cap_$TAG.txt
for cap in 16777216 33554432 67108864; do
  for wl in lexingtonHII20 clumpy256; do
    CMIB_QUEUE_CAPACITY=$cap timeout 600 python bench.py --workload $wl --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('$wl cap $cap: %.1f ms/step  %.3e packets/s  march %.1f prepare %.1f update %.1f rounds %.0f' % (d['ms_per_step'], d['value'], r['kernel_ms'], r['prepare_kernel_ms'], r['update_state_kernel_ms'], r['kernel_launches_per_step']))" >> $OUT/cap_$TAG.txt
  done
done
cat $OUT/cap_$TAG.txt
