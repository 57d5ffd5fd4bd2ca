#!/bin/bash
# round 2 / session 6: grouped DCN kernels, bench --config for BASELINE configs 3/4/5, full GPU suite
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dcn_fused.py -x -q -k "grouped" 2>&1 | tail -15
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for c in segm_r50 pose_x101dcn bbox_x101dcn_ms bbox_r50; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s06_bench_$c.json 2> gpurun_out/s06_bench_$c.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/s06_bench_$c.json').read().strip().splitlines()[-1])
    c = d['roofline']['classes']
    print('$c', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'serial', round(d['roofline']['serialized_step_ms'], 2),
          {k.split('(')[0]: round(v['ms_per_step'], 2) for k, v in c.items()}, 'launches', d['gpu_launches'], d.get('shapes'))
except Exception as e:
    print('$c', 'FAILED', e); print(open('gpurun_out/s06_bench_$c.err').read()[-2500:])
PY
done
