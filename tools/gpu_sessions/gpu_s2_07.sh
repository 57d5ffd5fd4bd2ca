#!/bin/bash
mkdir -p gpurun_out
for v in 28 34 44; do for pa in 0 2; do
LSNET_BIN_VARIANT=$v LSNET_BIN_PATCH=$pa timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_v.json').read().strip().splitlines()[-1])
    c=d['roofline']['classes']
    print('variant $v patch $pa: ms/step', round(d['ms_per_step'],2), 'col2im', round(c['dcn_col2im(scatter)']['ms_per_step'],2), 'wgrad', round(c['gemm_mnmajor(tcgen05 weight grad)']['ms_per_step'],2))
except Exception as e: print('parse fail', e)
PY
done; done
LSNET_OVERLAP_WGRAD=0 LSNET_BIN_VARIANT=44 LSNET_BIN_PATCH=2 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['roofline']['classes']
print('no wgrad overlap v44 p2: ms/step', round(d['ms_per_step'],2), 'col2im', round(c['dcn_col2im(scatter)']['ms_per_step'],2), 'wgrad', round(c['gemm_mnmajor(tcgen05 weight grad)']['ms_per_step'],2))"
