#!/bin/bash
# r02: task-based packet conventions on the device + the reference's task-based run of lexingtonHII20
TAG=${1:-tb}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_shoot.py tests/test_gpu_benchmarks.py -m gpu -q --timeout 900 -k "conventions or lexingtonHII20 or stromgren_diffuse" > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -30 $OUT/gpu_tests_$TAG.log | cut -c1-400
