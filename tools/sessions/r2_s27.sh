#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trunk_conv.py tests/test_gpu_dcn_fused.py -q -x 2>&1 | tail -3
timeout 300 python tools/bench_kernels.py 2>&1 | head -14
timeout 900 python tools/bench_trunk.py gpurun_out/r02_trunk_layers.md > gpurun_out/s27_trunk.log 2>&1; tail -27 gpurun_out/s27_trunk.log | cut -c1-100
for i in 1 2; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s27_bench_$i.json 2> gpurun_out/s27_bench_$i.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/s27_bench_$i.json').read().strip().splitlines()[-1])
c = d['roofline']['classes']
print('$i', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'serial', round(d['roofline']['serialized_step_ms'],2), {k.split('(')[0]: (round(v['ms_per_step'],2), round(v['achieved'])) for k, v in c.items()})
PY
done
