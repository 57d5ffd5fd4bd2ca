#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "dcn" -p no:cacheprovider > gpurun_out/pytest_dcn.log 2>&1
grep -E "^E  +(assert|Assert)|passed|failed|^FAILED" gpurun_out/pytest_dcn.log | head -12
for v in 28 38 34 44 48 54; do for pa in 1 2; do
echo "--- variant=$v patch=$pa: $(LSNET_BIN_VARIANT=$v LSNET_BIN_PATCH=$pa timeout 300 python tools/bench_kernels.py --only col2im 2>&1 | grep -E "col2im" | grep -oE '[0-9.]+ ms' | tr '\n' ' ')"
done; done
for sk in 1 2 3; do
echo "--- variant=44 patch=2 skip=$sk: $(LSNET_BIN_VARIANT=44 LSNET_BIN_PATCH=2 LSNET_BIN_SKIP=$sk timeout 300 python tools/bench_kernels.py --only col2im 2>&1 | grep -E "col2im" | grep -oE '[0-9.]+ ms' | tr '\n' ' ')"
done
LSNET_BIN_VARIANT=44 LSNET_BIN_PATCH=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dcn_col2im" -s 1 -c 1 -o gpurun_out/full_col2im_b44 -f python tools/bench_kernels.py --ncu col2im > gpurun_out/ncu_full_col2im_binned.log 2>&1; echo "ncu exit $?"
