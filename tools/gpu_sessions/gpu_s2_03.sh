#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1
grep -E "^E  +(assert|Assert)|passed|failed|^FAILED|Error" gpurun_out/pytest_model.log | head -12
for ls in 0 1; do
LSNET_LEVEL_STREAMS=$ls timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ls$ls.json 2> gpurun_out/bench_ls$ls.err; echo "bench ls=$ls exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_ls$ls.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
    for k,v in d['roofline']['classes'].items(): print(' ', k, v['ms_per_step'], v['achieved'])
except Exception as e: print('parse fail', e)
PY
tail -2 gpurun_out/bench_ls$ls.err
done
