#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py 2>&1 | tail -5
for grp in "conv2d" "dcn"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$grp" -rA -p no:cacheprovider > "gpurun_out/pytest_${name}.log" 2>&1
  echo "exit $?" >> "gpurun_out/pytest_${name}.log"
done
for f in gpurun_out/pytest_conv2d.log gpurun_out/pytest_dcn.log; do echo "=== $f"; grep -E "^(PASSED|FAILED|ERROR)|assert|Error|exit" "$f" | head -60; done
