#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_trajectory.py tests/test_gpu_dcn_fused.py -q -k "free_running or graph_cache or grouped" 2>&1 | tail -6
timeout 600 python tools/ref_kernel_table.py gpurun_out/r02_reference_kernels.md 2>&1 | grep "X-101"
for c in pose_x101dcn bbox_x101dcn_ms; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s08_bench_$c.json 2> gpurun_out/s08_bench_$c.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/s08_bench_$c.json').read().strip().splitlines()[-1])
    c = d['roofline']['classes']
    print('$c', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'serial', round(d['roofline']['serialized_step_ms'], 2),
          {k.split('(')[0]: round(v['ms_per_step'], 2) for k, v in c.items()})
except Exception as e:
    print('$c', 'FAILED', e); print(open('gpurun_out/s08_bench_$c.err').read()[-2500:])
PY
done
