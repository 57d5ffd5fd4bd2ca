#!/bin/bash
cd /root/repo
CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/sessions/dbg_capture.py 2>&1 | grep -v Warning | tail -40
timeout 600 python -m pytest tests/test_gpu_trajectory.py -x -q -k prefetch 2>&1 | grep -v Warning | tail -30
