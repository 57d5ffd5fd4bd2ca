#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kmajor -c 2 -o gpurun_out/r02_trunk_l1_64_256_v2 -f python tools/bench_trunk.py /tmp/x.md --only "l1 1x1 64>256" > gpurun_out/s36_ncu1.log 2>&1; echo "ncu1 $?"
