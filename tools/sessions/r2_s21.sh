#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trunk_conv.py -x -q -k "own_convs_match or fusion_is_exact" 2>&1 | grep -v Warn | tail -40
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "pack_fused_sigmoid" 2>&1 | grep -v Warn | tail -40
timeout 600 python -m pytest tests/test_gpu_dcn_fused.py -x -q -k "gradient_sink" 2>&1 | grep -v Warn | tail -30
