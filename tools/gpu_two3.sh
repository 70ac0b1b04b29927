#!/bin/bash
TAG=${1:-two3}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_comm.py tests/test_gpu_host_driver.py -m gpu -q --timeout 900 > $OUT/gpu_tests_$TAG.log 2>&1
echo "pytest rc=$?"; tail -5 $OUT/gpu_tests_$TAG.log | cut -c1-400
grep -h "AssertionError\|Error:" $OUT/comm_test_*.err | head -10
