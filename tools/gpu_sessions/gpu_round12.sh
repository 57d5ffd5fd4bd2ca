#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "dcn_forward" -p no:cacheprovider 2>&1 | tail -2
for ov in 0 1; do
LSNET_OVERLAP_WGRAD=$ov timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ov$ov.json 2> gpurun_out/bench_ov$ov.err; echo "bench overlap=$ov exit $?"
python - $ov <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/bench_ov%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
except Exception as e: print('parse fail', e)
PY
tail -2 gpurun_out/bench_ov$ov.err
done
