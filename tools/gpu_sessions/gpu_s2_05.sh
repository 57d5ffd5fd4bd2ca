#!/bin/bash
mkdir -p gpurun_out
for sk in 0 1 2; do
LSNET_BIN_SKIP=$sk LSNET_BIN_VARIANT=40 LSNET_BIN_PATCH=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dcn_col2im" -s 1 -c 1 -o gpurun_out/full_col2im_b40_sk$sk -f python tools/bench_kernels.py --ncu col2im > gpurun_out/ncu_full_col2im_binned.log 2>&1; echo "ncu exit $?"
done
ls -la gpurun_out/*.ncu-rep
