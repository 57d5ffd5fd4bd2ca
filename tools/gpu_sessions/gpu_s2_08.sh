#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "dcn" -p no:cacheprovider > gpurun_out/pytest_dcn.log 2>&1
grep -E "^E  +|passed|failed|^FAILED" gpurun_out/pytest_dcn.log | head -20
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "resnext" -p no:cacheprovider > gpurun_out/pytest_x101.log 2>&1
grep -E "^E  +|passed|failed|^FAILED|Error" gpurun_out/pytest_x101.log | head -30
timeout 300 python tools/bench_kernels.py --only col2im 2>&1 | grep -E "col2im"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
    for k,v in d['roofline']['classes'].items(): print(' ', k, v['ms_per_step'], v['achieved'])
except Exception as e: print('parse fail', e)
PY
