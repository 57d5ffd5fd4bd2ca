#!/bin/bash
# round 2 / session 12: current timeline (CUPTI), per-class DRAM traffic of one step, per-kernel bench
cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/trace_step.py > gpurun_out/s12_trace.log 2>&1; echo "trace exit $?"; head -5 gpurun_out/trace_summary.md
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_traffic.csv python tools/profile_step.py > gpurun_out/s12_ncu_step.log 2>&1; echo "ncu step exit $?"; wc -l gpurun_out/step_traffic.csv
timeout 600 python tools/bench_kernels.py > gpurun_out/s12_bench_kernels.log 2>&1; echo "bench_kernels exit $?"; tail -30 gpurun_out/s12_bench_kernels.log
