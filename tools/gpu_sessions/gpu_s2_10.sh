#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
grep -E "^E  +|passed|failed|^FAILED" gpurun_out/pytest_gpu.log | cut -c1-220 | head -30
for pk in 1 0; do
LSNET_DCN_PACKED_OM=$pk timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pk$pk.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_pk$pk.json').read().strip().splitlines()[-1])
    print('packed=$pk value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
except Exception as e: print('parse fail', e)
PY
done
