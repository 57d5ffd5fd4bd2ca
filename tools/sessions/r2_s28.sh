#!/bin/bash
cd /root/repo
for v in cur prev s14; do
  if [ $v = cur ]; then unset LSNET_LIB_PATH; else export LSNET_LIB_PATH=/root/repo/tools/sessions/_alt/liblsnet_$v.so; fi
  echo "== $v"; timeout 300 python tools/bench_kernels.py 2>&1 | head -5
done
