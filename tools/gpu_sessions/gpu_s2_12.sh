#!/bin/bash
for rep in 1 2; do for pk in 1 0; do
LSNET_DCN_PACKED_OM=$pk timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['roofline']['classes']
print('packed=$pk ms/step', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'col2im', round(c['dcn_col2im(scatter)']['ms_per_step'],2), 'im2col', round(c['dcn_im2col(gather)']['ms_per_step'],2), 'clk', d['clocks'])"
done; done
