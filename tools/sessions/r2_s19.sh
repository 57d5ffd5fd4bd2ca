#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "detector_vs_oracle_bbox" -s 2>&1 | grep -v Warning | tail -30
LSNET_TRUNK=cudnn timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "detector_vs_oracle_bbox" -s 2>&1 | grep -v Warning | tail -12
LSNET_STEM_OWN=0 timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "detector_vs_oracle_bbox" -s 2>&1 | grep -v Warning | tail -12
