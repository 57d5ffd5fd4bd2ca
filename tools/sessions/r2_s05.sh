#!/bin/bash
# round 2 / session 5: whole step with the TMA-staged adjoint, full GPU suite
cd /root/repo
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s05_bench_$name.json 2> gpurun_out/s05_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/s05_bench_$name.json').read().strip().splitlines()[-1])
    c = d['roofline']['classes']
    print('$name', 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 2), 'serial', round(d['roofline']['serialized_step_ms'], 2),
          {k.split('(')[0]: round(v['ms_per_step'], 2) for k, v in c.items()}, 'launches', d['gpu_launches'])
except Exception as e:
    print('$name', 'FAILED', e); print(open('gpurun_out/s05_bench_$name.err').read()[-1500:])
PY
}
run default
run adj_off LSNET_ADJOINT_TMA=0
run adj_all LSNET_ADJOINT_TMA=2
run unfused_adj LSNET_DCN_FUSED=0
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
