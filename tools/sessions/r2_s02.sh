#!/bin/bash
# round 2 / session 2: pipelined fused forward + fused weight gradient
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dcn_fused.py -x -q 2>&1 | tail -15 > gpurun_out/s02_fused_tests.log
echo "fused tests exit ${PIPESTATUS[0]}" >> gpurun_out/s02_fused_tests.log
for st in 2 3; do
  LSNET_DCN_FUSED_STAGES=$st timeout 600 python tools/bench_kernels.py --only dcn_fwd 2>&1 | head -12 > gpurun_out/s02_bench_fwd_st$st.log
done
timeout 600 python tools/bench_kernels.py --only dcn_wgrad > gpurun_out/s02_bench_wgrad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcn_fused_fwd -s 6 -c 1 -o gpurun_out/s02_fused_fwd -f python tools/bench_kernels.py --ncu dcn_fwd > gpurun_out/s02_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcn_fused_wgrad -s 2 -c 1 -o gpurun_out/s02_fused_wgrad -f python tools/bench_kernels.py --ncu dcn_wgrad >> gpurun_out/s02_ncu.log 2>&1
cat gpurun_out/s02_fused_tests.log | tail -8
cat gpurun_out/s02_bench_fwd_st2.log gpurun_out/s02_bench_wgrad.log
