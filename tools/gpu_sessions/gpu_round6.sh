#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_model.py -q -m gpu -rA -s -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1; echo "exit $?" >> gpurun_out/pytest_model.log
grep -E "^(FAILED|ERROR)|^E    |passed|failed|exit" gpurun_out/pytest_model.log | head -30
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "dcn" -p no:cacheprovider 2>&1 | tail -3
for u in 1 2 3; do echo "im2col U=$u"; LSNET_IM2COL_U=$u timeout 300 python tools/bench_kernels.py 2>&1 | grep -E "im2col|col2im"; done
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['config']['step_mode'])
    for k,v in d['roofline']['classes'].items(): print(' ', k, v)
except Exception as e: print('parse fail', e)
PY
tail -n 12 gpurun_out/bench.err
