#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s11_bench_$name.json 2> gpurun_out/s11_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/s11_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e img/s', round(d['e2e']['value'], 1), 'e2e ms', round(d['e2e']['ms_per_step'], 2), 'serial', round(d['roofline']['serialized_step_ms'], 2), 'launches', d['gpu_launches'])
except Exception as e:
    print('$name', 'FAILED', e); print(open('gpurun_out/s11_bench_$name.err').read()[-1500:])
PY
}
run default
run novec LSNET_DIRECT_VEC=0
run default2
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
