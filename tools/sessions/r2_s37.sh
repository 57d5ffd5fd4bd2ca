#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s37_bench_$name.json 2> gpurun_out/s37_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/s37_bench_$name.json').read().strip().splitlines()[-1])
    c = d['roofline']['classes']
    print('$name', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'serial', round(d['roofline']['serialized_step_ms'],2), {k.split('(')[0]: (round(v['ms_per_step'],2), round(v['achieved'])) for k, v in c.items() if 'gemm_k' in k})
except Exception as e:
    print('$name', 'FAILED', e); print(open('gpurun_out/s37_bench_$name.err').read()[-800:])
PY
}
run a
run adapt2 LSNET_GEMM_ADAPT_BN=2
run b
run adapt2b LSNET_GEMM_ADAPT_BN=2
