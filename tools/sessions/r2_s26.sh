#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for i in 1 2 3 4; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s26_bench_$i.json 2> gpurun_out/s26_bench_$i.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/s26_bench_$i.json').read().strip().splitlines()[-1])
print('$i', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), round(d['e2e']['ms_per_step'],2), d['e2e']['step_wall_ms'])
PY
done
