#!/bin/bash
# round 2 final evidence: full gpu tests, bench lines of all configs, A/B of the late knobs, per-class DRAM traffic, timeline
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/fe_bench_$name.json 2> gpurun_out/fe_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/fe_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'serial', round(d['roofline']['serialized_step_ms'],2))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/fe_bench_$name.err').read()[-800:])
PY
}
run a
run nosink LSNET_BIAS_SINK=0
run b
run nosink2 LSNET_BIAS_SINK=0
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/fe_bench_full.json 2> gpurun_out/fe_bench_full.err; tail -c 600 gpurun_out/fe_bench_full.json
timeout 600 python tools/trace_step.py > gpurun_out/fe_trace.log 2>&1; head -4 gpurun_out/trace_summary.md
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_traffic.csv python tools/profile_step.py > gpurun_out/fe_ncu_step.log 2>&1; echo "ncu step exit $?"; wc -l gpurun_out/step_traffic.csv
