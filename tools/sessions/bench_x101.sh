#!/bin/bash
# round 2, last build: the two X-101-64x4d-DCN configs (multi-scale bbox, 17-keypoint pose) through bench.py
cd /root/repo
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 420 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/fin_bench_$name.json 2> gpurun_out/fin_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/fin_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'serial', round(d['roofline']['serialized_step_ms'],2), d['roofline']['kernel'][:24], round(d['roofline']['frac'],3), d['shapes'])
except Exception as e:
    print('$name', 'FAILED', e); print(open('gpurun_out/fin_bench_$name.err').read()[-1200:])
PY
}
run pose_x101dcn --config pose_x101dcn
run bbox_x101dcn_ms --config bbox_x101dcn_ms
