#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench1.json; tail -n 15 gpurun_out/bench1.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu exit $?"
tail -n 3 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches.csv
