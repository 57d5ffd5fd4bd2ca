#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm or conv" -p no:cacheprovider 2>&1 | tail -2
for ad in 1 0; do
LSNET_GEMM_ADAPT_BN=$ad timeout 600 python tools/trace_step.py > /dev/null 2>&1
echo "adapt=$ad $(sed -n 3p gpurun_out/trace_summary.md) $(grep gemm_kmajor gpurun_out/trace_summary.md)"
done
