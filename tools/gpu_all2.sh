#!/bin/bash
# r02, 2 GPUs: the whole GPU tier (2-GPU tests included)
TAG=${1:-all2}
OUT=gpurun_out
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 --durations=15 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -30 $OUT/gpu_tests_$TAG.log | cut -c1-300
grep -h "AssertionError\|Error:" $OUT/comm_test_*.err 2>/dev/null | head -10
