#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
grep -E "^E  +|passed|failed|^FAILED" gpurun_out/pytest_gpu.log | cut -c1-220 | head -30
for dw in 1 0; do
LSNET_DIRECT_WGRAD=$dw timeout 600 python tools/trace_step.py > /dev/null 2>&1
echo "direct=$dw $(sed -n 3p gpurun_out/trace_summary.md)"
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['roofline']['classes']
print('ms/step', round(d['ms_per_step'],2), 'img/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
