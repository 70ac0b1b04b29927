#!/bin/bash
# N-GPU strong-scaling line as the driver launches it: headline workload + the 256^3 north-star grids under `workloads`
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
N=${2:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_${N}gpu_$TAG.json 2> $OUT/bench_${N}gpu_$TAG.err
echo "bench $N rc=$?"; tail -3 $OUT/bench_${N}gpu_$TAG.err | cut -c1-300
python - <<P
import json
d = json.loads(open("$OUT/bench_${N}gpu_$TAG.json").read().strip().splitlines()[-1])
def show(name, r):
    print(name, "%.2f ms/step %.3e packets/s" % (r["ms_per_step"], r["value"]), "e2e", r["e2e"] and "%.3e" % r["e2e"]["value"],
          "roof %s %.3f" % (r["roofline"]["bound"], r["roofline"]["frac"]), {k: round(v, 2) for k, v in r["phases_ms"].items()})
show("head", d)
for k, r in d.get("workloads", {}).items(): show("   " + k, r)
if "weak" in d: print("   weak", d["weak"]["value"], d["weak"]["ms_per_step"])
P
