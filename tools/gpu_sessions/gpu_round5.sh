#!/bin/bash
mkdir -p gpurun_out
for grp in "gemm" "conv2d" "dcn_forward"; do
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$grp" -rA -p no:cacheprovider > gpurun_out/pytest_$grp.log 2>&1; echo "exit $?" >> gpurun_out/pytest_$grp.log
grep -E "^(FAILED|ERROR)|^E    |passed|failed|exit" gpurun_out/pytest_$grp.log | head -12
done
timeout 1200 python -m pytest tests/test_gpu_model.py -q -m gpu -rA -s -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1; echo "exit $?" >> gpurun_out/pytest_model.log
grep -E "^(FAILED|ERROR)|^E    |passed|failed|exit" gpurun_out/pytest_model.log | head -30
timeout 600 python tools/bench_kernels.py 2>&1 | tail -12
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['config']['step_mode'])
    for k,v in d['roofline']['classes'].items(): print(' ', k, v)
    print({k:v for k,v in d['roofline'].items() if k!='classes'})
except Exception as e: print('parse fail', e)
PY
tail -n 12 gpurun_out/bench.err
