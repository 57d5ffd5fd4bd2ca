#!/bin/bash
for pk in 1 0; do
LSNET_DCN_PACKED_OM=$pk timeout 600 python tools/trace_step.py > /dev/null 2>&1
cp gpurun_out/trace_summary.md gpurun_out/trace_summary_pk$pk.md
done
