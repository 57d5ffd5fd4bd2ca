#!/bin/bash
# round 2 / session 1: fused DCN forward — parity, regression, micro-benchmarks, one ncu capture
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s01_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_dcn_fused.py -x -q -rA 2>&1 | tail -60 > gpurun_out/s01_fused_tests.log
echo "fused tests exit ${PIPESTATUS[0]}" >> gpurun_out/s01_fused_tests.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "dcn" 2>&1 | tail -15 > gpurun_out/s01_dcn_regress.log
for st in 2 3; do
  LSNET_DCN_FUSED_STAGES=$st timeout 600 python tools/bench_kernels.py --only dcn_fwd > gpurun_out/s01_bench_st$st.log 2>&1
done
timeout 300 python tools/bench_kernels.py --only im2col >> gpurun_out/s01_bench_st2.log 2>&1
timeout 300 python tools/bench_kernels.py --only gemm >> gpurun_out/s01_bench_st2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcn_fused_fwd -s 3 -c 1 -o gpurun_out/s01_fused_fwd -f python tools/bench_kernels.py --ncu dcn_fwd > gpurun_out/s01_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/s01_ncu.log
tail -5 gpurun_out/s01_fused_tests.log
cat gpurun_out/s01_bench_st2.log | tail -40
