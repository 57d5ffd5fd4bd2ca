#!/bin/bash
# first GPU contact: diag + grouped pytest runs (separate processes so one CUDA fault does not poison the rest)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python tools/gpu_diag.py > gpurun_out/diag.log 2>&1; echo "diag exit $?" >> gpurun_out/diag.log
for grp in "gemm" "conv2d" "dcn" "cross_iou or focal or directional" "assign or pred_boxes"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$grp" -rA -p no:cacheprovider > "gpurun_out/pytest_${name}.log" 2>&1
  echo "exit $?" >> "gpurun_out/pytest_${name}.log"
done
tail -n 30 gpurun_out/diag.log
for f in gpurun_out/pytest_*.log; do echo "=== $f"; tail -n 25 "$f"; done
