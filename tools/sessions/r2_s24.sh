#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "detector_vs_oracle_bbox" -s 2>&1 | grep "COS\|passed\|failed" | head -20
LSNET_STEM_OWN=0 timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "detector_vs_oracle_bbox" -s 2>&1 | grep "COS\|passed\|failed" | head -20
LSNET_GX_SINK=0 LSNET_TRUNK_BWD_FUSE=0 timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "detector_vs_oracle_bbox" -s 2>&1 | grep "COS\|passed\|failed" | head -20
timeout 600 python -m pytest tests/test_gpu_compat_ext.py -q 2>&1 | tail -15
