#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_diag2.py 2>&1 | grep -v Warning | tail -12
for m in 0 1; do
echo "== detector test DX_FP32=$m"
LSNET_DCN_DX_FP32=$m timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "detector_vs_oracle" -rA -s -p no:cacheprovider 2>&1 | grep -E "COS|passed|failed|^E  +Assert" | head -12
done
for grp in "gemm" "conv2d"; do
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$grp" -p no:cacheprovider 2>&1 | tail -2
done
timeout 300 python tools/bench_kernels.py 2>&1 | grep -E "gemm"
