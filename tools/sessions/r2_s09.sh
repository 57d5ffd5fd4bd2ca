#!/bin/bash
# round 2 / session 9: in-kernel bias/affine gradient accumulation + copy-stream prefetch: tests, step A/B, ncu captures of
# the kernels that had no summary yet (taken inside one real training step)
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s09_bench_$name.json 2> gpurun_out/s09_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/s09_bench_$name.json').read().strip().splitlines()[-1])
    c = d['roofline']['classes']
    print('$name', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e img/s', round(d['e2e']['value'], 1), 'e2e ms', round(d['e2e']['ms_per_step'], 2), 'serial', round(d['roofline']['serialized_step_ms'], 2), 'traffic', d['roofline']['traffic'])
except Exception as e:
    print('$name', 'FAILED', e); print(open('gpurun_out/s09_bench_$name.err').read()[-1500:])
PY
}
run default
run novec LSNET_DIRECT_VEC=0
for k in cross_iou_fused focal_fwd focal_bwd atss_candidates gn_stats gn_apply gn_bwd_stats gn_bwd_apply grad_prep_vec; do
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"^$k|::$k" -c 1 -o gpurun_out/r02_step_$k -f python tools/profile_step.py > gpurun_out/s09_ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"cross_iou_fused" -c 1 -o gpurun_out/r02_step_cross_iou_fused_segm -f python tools/profile_step.py --config segm_r50 > gpurun_out/s09_ncu_segm.log 2>&1; echo "ncu segm exit $?"
ls -la gpurun_out/r02_step_*.ncu-rep | awk '{print $5, $9}'
