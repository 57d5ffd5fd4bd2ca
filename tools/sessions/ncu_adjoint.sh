#!/bin/bash
# ncu --set full of the TMA-staged DCN adjoint (FHFMA build) at the level-0 shape
cd /root/repo
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dcn_adjoint_tma -s 1 -c 1 -o gpurun_out/r02_full_adjoint_fhfma -f python tools/bench_kernels.py --ncu col2im > gpurun_out/ncu_adjoint.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/r02_full_adjoint_fhfma.ncu-rep | awk '{print $5, $9}'
