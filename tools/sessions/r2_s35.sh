#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trunk_conv.py -q -x 2>&1 | tail -3
timeout 900 python tools/bench_trunk.py gpurun_out/trunk_layers_occ2.md > gpurun_out/s35_trunk.log 2>&1; tail -27 gpurun_out/s35_trunk.log | cut -c1-60 | head -24
for v in 1 0 1 0; do
  LSNET_GEMM_OCC2=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s35_bench_$v.json 2> gpurun_out/s35_bench_$v.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/s35_bench_$v.json').read().strip().splitlines()[-1])
c = d['roofline']['classes']
print('occ2=$v', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'serial', round(d['roofline']['serialized_step_ms'],2), {k.split('(')[0]: (round(v['ms_per_step'],2), round(v['achieved'])) for k, v in c.items() if 'gemm_k' in k})
PY
done
