#!/bin/bash
# round 2 / session 4: TMA-staged adjoint — parity, timing vs the binned kernel, ncu
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dcn_fused.py tests/test_gpu_kernels.py -x -q -k "dcn or adjoint or col2im" 2>&1 | tail -12 > gpurun_out/s04_tests.log
echo "tests exit ${PIPESTATUS[0]}" >> gpurun_out/s04_tests.log
for v in 1 0; do
  LSNET_ADJOINT_TMA=$v timeout 600 python tools/bench_kernels.py --only col2im > gpurun_out/s04_col2im_tma$v.log 2>&1
  LSNET_ADJOINT_TMA=$v timeout 600 python tools/bench_kernels.py --only col2im_pyr >> gpurun_out/s04_col2im_tma$v.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcn_adjoint_tma -s 3 -c 1 -o gpurun_out/s04_adjoint -f python tools/bench_kernels.py --ncu col2im > gpurun_out/s04_ncu.log 2>&1
cat gpurun_out/s04_tests.log | tail -6
echo "--- TMA adjoint"; cat gpurun_out/s04_col2im_tma1.log
echo "--- binned (r01)"; cat gpurun_out/s04_col2im_tma0.log
