#!/bin/bash
# round 2 / session 7: trajectory / multi-scale parity tests, per-parameter table, reference kernel table, step traffic
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_trajectory.py tests/test_gpu_decode.py -q -rA 2>&1 | grep -v "Warning\|warn" | tail -40
timeout 900 python tools/param_grad_table.py gpurun_out/r02_param_grad_errors.md > gpurun_out/s07_param_table.log 2>&1; tail -3 gpurun_out/s07_param_table.log
timeout 900 python tools/ref_kernel_table.py gpurun_out/r02_reference_kernels.md > gpurun_out/s07_ref_kernels.log 2>&1; tail -14 gpurun_out/s07_ref_kernels.log
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_traffic.csv python tools/profile_step.py > gpurun_out/s07_ncu_step.log 2>&1; echo "ncu step exit $?"; wc -l gpurun_out/step_traffic.csv
