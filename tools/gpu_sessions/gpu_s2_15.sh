#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "graph_trainer or oracle_bbox" -p no:cacheprovider 2>&1 | tail -2
for grp in "0,1|2|3|4" "0|1|2|3|4" "0,1|2,3,4"; do
LSNET_LEVEL_STREAMS=1 LSNET_LEVEL_GROUPS="$grp" timeout 600 python tools/trace_step.py > /dev/null 2>&1
echo "groups=$grp $(sed -n 3p gpurun_out/trace_summary.md)"
LSNET_LEVEL_STREAMS=1 LSNET_LEVEL_GROUPS="$grp" timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  bench ms/step', round(d['ms_per_step'],2), 'img/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
done
