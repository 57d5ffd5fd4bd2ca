#!/bin/bash
# round 2: input-side check on the GPU -- new data-path tests, whole gpu suite, e2e with float vs uint8 host batches,
# ncu --set full of lsnet_image_prep_u8 at the bench canvas (B4 800x1344)
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_datapath.py -q -x > gpurun_out/dp_tests.log 2>&1; tail -5 gpurun_out/dp_tests.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/dp_suite.log 2>&1; tail -3 gpurun_out/dp_suite.log
for inp in u8 f32 u8; do
  n=$(ls gpurun_out/dp_bench_${inp}_*.json 2>/dev/null | wc -l)
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-input $inp > gpurun_out/dp_bench_${inp}_$n.json 2> gpurun_out/dp_bench_${inp}_$n.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/dp_bench_${inp}_$n.json').read().strip().splitlines()[-1])
    print('$inp', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'h2d', d['e2e']['h2d_bytes_per_step'], d['e2e']['step_wall_ms'])
except Exception as e:
    print('$inp FAILED', e); print(open('gpurun_out/dp_bench_${inp}_$n.err').read()[-1200:])
PY
done
timeout 120 python tools/bench_image_prep.py
timeout 300 ncu --set full --clock-control none --import-source on -k regex:image_prep -s 1 -c 1 -o gpurun_out/r02_full_image_prep -f python tools/bench_image_prep.py > gpurun_out/dp_ncu_prep.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/r02_full_image_prep.ncu-rep | awk '{print $5, $9}'
