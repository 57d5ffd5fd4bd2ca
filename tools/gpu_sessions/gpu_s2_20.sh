#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_graph.csv python tools/profile_step.py > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
tail -n 1 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches_graph.csv
timeout 600 python tools/trace_step.py > /dev/null 2>&1; sed -n 3p gpurun_out/trace_summary.md
