#!/bin/bash
# round 2 / session 13: own trunk convolutions (strided implicit GEMM, phase dgrad, fold packs, stem), split-K quantisation fix
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trunk_conv.py -q 2>&1 | grep -v "Warning\|warn" > gpurun_out/s13_tests.log; tail -15 gpurun_out/s13_tests.log
timeout 300 python tools/bench_kernels.py 2>&1 | grep -i "gemm_mnmajor\|backward_weight saved" | head
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s13_bench_$name.json 2> gpurun_out/s13_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/s13_bench_$name.json').read().strip().splitlines()[-1])
    print('$name', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e img/s', round(d['e2e']['value'], 1), 'serial', round(d['roofline']['serialized_step_ms'], 2), 'launches', d['gpu_launches'])
except Exception as e:
    print('$name', 'FAILED', e); print(open('gpurun_out/s13_bench_$name.err').read()[-800:])
PY
}
run own
run cudnn LSNET_TRUNK=cudnn
run own_cudnnstem LSNET_STEM_OWN=0
run own_mt2 LSNET_WGRAD_MT2=1
run own_nofoldside LSNET_FOLD_SIDE=0
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
