#!/bin/bash
# r02, 2 GPUs: the comm ABI test, the C++ 2-GPU driver test, bench.py at N=2 (short)
TAG=${1:-two}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_comm.py tests/test_gpu_host_driver.py -m gpu -q -x --timeout 900 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -15 $OUT/gpu_tests_$TAG.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/bench_2gpu_$TAG.json 2> $OUT/bench_2gpu_$TAG.err
echo "bench2 rc=$?"; tail -3 $OUT/bench_2gpu_$TAG.err | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench_1gpu_$TAG.json 2> $OUT/bench_1gpu_$TAG.err
echo "bench1 rc=$?"; tail -3 $OUT/bench_1gpu_$TAG.err | cut -c1-300
python - <<P
import json
for f in ("$OUT/bench_1gpu_$TAG.json", "$OUT/bench_2gpu_$TAG.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    def show(name, r):
        print(name, "%.1f ms/step %.3e packets/s" % (r["ms_per_step"], r["value"]), "e2e", r["e2e"] and "%.3e" % r["e2e"]["value"],
              "roof %s %.3f" % (r["roofline"]["bound"], r["roofline"]["frac"]), {k: round(v, 2) for k, v in r["phases_ms"].items()})
    show(f + " head", d)
    for k, r in d.get("workloads", {}).items(): show("   " + k, r)
    if "weak" in d: print("   weak", d["weak"]["value"], d["weak"]["ms_per_step"])
P
