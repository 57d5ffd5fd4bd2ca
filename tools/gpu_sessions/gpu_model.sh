#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_model.py -q -m gpu -rA -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1
echo "exit $?" >> gpurun_out/pytest_model.log
grep -E "^(PASSED|FAILED|ERROR)|assert|Error|exit|^E " gpurun_out/pytest_model.log | head -60
