#!/bin/bash
timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | grep -E "^E   +|passed|failed|^FAILED" | cut -c1-200
for hg in 1 0; do LSNET_HEAD_GLUE=$hg timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('glue=$hg ms/step', round(d['ms_per_step'],2), 'img/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"; done
