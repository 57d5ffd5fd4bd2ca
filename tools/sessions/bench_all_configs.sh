#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
run() { name=$1; shift
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/s30_bench_$name.json 2> gpurun_out/s30_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/s30_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'serial', round(d['roofline']['serialized_step_ms'],2), d['roofline']['kernel'][:24], round(d['roofline']['frac'],3))
except Exception as e:
    print('$name', 'FAILED', e); print(open('gpurun_out/s30_bench_$name.err').read()[-1200:])
PY
}
run bbox_r50 --steps 20 --warmup 5
run segm_r50 --config segm_r50
run bbox_x101dcn_ms --config bbox_x101dcn_ms
run pose_x101dcn --config pose_x101dcn
