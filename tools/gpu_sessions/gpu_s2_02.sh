#!/bin/bash
# time split of the binned col2im: skip phases / patch sizes
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "dcn" -p no:cacheprovider > gpurun_out/pytest_dcn.log 2>&1
grep -E "^E  +(assert|Assert)|passed|failed|^FAILED" gpurun_out/pytest_dcn.log | head -12
for pa in 0 1 2; do
echo "--- skip=0 patch=$pa"; LSNET_BIN_SKIP=0 LSNET_BIN_PATCH=$pa timeout 300 python tools/bench_kernels.py --only col2im 2>&1 | grep -E "col2im"
done
for sk in 1 2; do 
echo "--- skip=$sk patch=2"; LSNET_BIN_SKIP=$sk LSNET_BIN_PATCH=2 timeout 300 python tools/bench_kernels.py --only col2im 2>&1 | grep -E "col2im"
done
echo "--- pyr patch=2";  LSNET_BIN_PATCH=2 timeout 300 python tools/bench_kernels.py --only col2im_pyr 2>&1 | grep -E "col2im"
