#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "dcn" -p no:cacheprovider > gpurun_out/pytest_dcn.log 2>&1
grep -E "^E  +|passed|failed|^FAILED" gpurun_out/pytest_dcn.log | head -20
for u in 1 2 3; do echo "U=$u $(LSNET_IM2COL_U=$u timeout 300 python tools/bench_kernels.py --only im2col 2>&1 | grep -E "im2col")"; done
LSNET_IM2COL_U=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dcn_im2col" -s 1 -c 1 -o gpurun_out/full_im2col_v2 -f python tools/bench_kernels.py --ncu im2col > gpurun_out/ncu_full_im2col.log 2>&1; echo "ncu exit $?"
