#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_graph.csv python tools/profile_step.py > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
tail -n 1 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches_graph.csv
for k in gemm col2im im2col wgrad conv; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_kmajor|gemm_mnmajor|dcn_col2im|dcn_im2col" -s 1 -c 1 -o gpurun_out/full_$k -f python tools/bench_kernels.py --ncu $k > gpurun_out/ncu_full_$k.log 2>&1; echo "ncu full $k exit $?"
done
ls -la gpurun_out/*.ncu-rep
