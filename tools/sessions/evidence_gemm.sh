#!/bin/bash
# round 2: ncu --set full of the two headline GEMM kernels at the level-0 shape after this round's changes
cd /root/repo
mkdir -p gpurun_out
for k in conv wgrad gemm; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_kmajor|gemm_mnmajor" -s 1 -c 1 -o gpurun_out/r02_full_$k -f python tools/bench_kernels.py --ncu $k > gpurun_out/ev_ncu_$k.log 2>&1; echo "ncu full $k exit $?"
done
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ls -la gpurun_out/r02_full_*.ncu-rep | awk '{print $5, $9}'
