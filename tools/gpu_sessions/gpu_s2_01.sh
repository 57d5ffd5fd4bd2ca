#!/bin/bash
# session 2, call 1: binned col2im — parity, kernel timing (binned vs direct), whole step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "dcn" -p no:cacheprovider > gpurun_out/pytest_dcn.log 2>&1
grep -E "^E  +(assert|Assert)|passed|failed|^FAILED" gpurun_out/pytest_dcn.log | head -12
echo "--- binned"; timeout 300 python tools/bench_kernels.py 2>&1 | grep -E "col2im|im2col"
cp gpurun_out/kernels.json gpurun_out/kernels_binned.json
echo "--- direct"; LSNET_COL2IM=direct timeout 300 python tools/bench_kernels.py 2>&1 | grep -E "col2im"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
    for k,v in d['roofline']['classes'].items(): print(' ', k, v)
except Exception as e: print('parse fail', e)
PY
tail -3 gpurun_out/bench.err
