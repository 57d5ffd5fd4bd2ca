#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_all_gpu.log 2>&1; echo "pytest -m gpu exit $?"
grep -E "^E  +(assert|Assert)|passed|failed|^FAILED|^ERROR" gpurun_out/pytest_all_gpu.log | head -12
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/bench_kernels.py 2>&1 | grep -E "groupnorm"
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
    print('roofline', {k:v for k,v in d['roofline'].items() if k!='classes'})
    for k,v in d['roofline']['classes'].items(): print(' ', k, v)
    print('cpu', d['cpu_baseline']); print(d['clocks'])
except Exception as e: print('parse fail', e)
PY
