#!/bin/bash
timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | grep -E "^E   +|passed|failed"
for i in 1 2; do timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],2), 'img/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"; done
timeout 600 python tools/trace_step.py > /dev/null 2>&1; sed -n 3p gpurun_out/trace_summary.md; grep "grad_prep" gpurun_out/trace_summary.md
