#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for v in 1 0 1; do
  LSNET_GN_EPILOGUE=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s41_bench_$v.json 2> gpurun_out/s41_bench_$v.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/s41_bench_$v.json').read().strip().splitlines()[-1])
    print('gn_epi=$v', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'serial', round(d['roofline']['serialized_step_ms'],2))
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/s41_bench_$v.err').read()[-800:])
PY
done
