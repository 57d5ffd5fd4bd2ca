#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "conv2d or dcn_forward" -p no:cacheprovider > gpurun_out/pytest_k.log 2>&1
grep -E "^E  +(assert|Assert)|passed|failed|^FAILED" gpurun_out/pytest_k.log | head -6
timeout 1200 python -m pytest tests/test_gpu_model.py -q -m gpu -rA -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1
grep -E "^E  +(assert|Assert)|passed|failed|^FAILED" gpurun_out/pytest_model.log | head -12
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
    for k,v in d['roofline']['classes'].items(): print(' ', k, v)
except Exception as e: print('parse fail', e)
PY
