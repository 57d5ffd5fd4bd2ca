#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s18_bench_$name.json 2> gpurun_out/s18_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/s18_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e img/s', round(d['e2e']['value'], 1), 'serial', round(d['roofline']['serialized_step_ms'], 2), 'launches', d['gpu_launches'])
except Exception as e:
    print('$name', 'FAILED', e); print(open('gpurun_out/s18_bench_$name.err').read()[-800:])
PY
}
run own
run noadapt LSNET_GEMM_ADAPT_BN=0
timeout 600 python tools/trace_step.py > gpurun_out/s18_trace.log 2>&1; echo "trace exit $?"; head -4 gpurun_out/trace_summary.md
