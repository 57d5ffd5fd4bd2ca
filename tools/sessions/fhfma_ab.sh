#!/bin/bash
# round 2: A/B of the adjoint's dOffset/dMask dot products on FHFMA.BF16 (default build) vs unpack + FFMA2
# (lsnet_b200/liblsnet_sm100_nofhfma.so, built with -DLSN_FHFMA=0): kernel timing, DCN parity tests, whole step
cd /root/repo
mkdir -p gpurun_out
OLD=/root/repo/lsnet_b200/liblsnet_sm100_nofhfma.so
echo "== adjoint micro-benchmark, FHFMA build"; timeout 200 python tools/bench_kernels.py --only col2im 2>&1 | grep -i "col2im" | head -5
echo "== adjoint micro-benchmark, FFMA2 build"; LSNET_LIB_PATH=$OLD timeout 200 python tools/bench_kernels.py --only col2im 2>&1 | grep -i "col2im" | head -5
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dcn_fused.py tests/test_gpu_compat_ext.py tests/test_gpu_reference_cuda.py -q -x > gpurun_out/fh_tests.log 2>&1; tail -3 gpurun_out/fh_tests.log
timeout 600 python -m pytest tests/test_gpu_datapath.py -q -x -s > gpurun_out/fh_datapath.log 2>&1; grep "loss:" gpurun_out/fh_datapath.log; tail -3 gpurun_out/fh_datapath.log
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/fh_bench_$name.json 2> gpurun_out/fh_bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/fh_bench_$name.json').read().strip().splitlines()[-1])
    c = d['roofline']['classes']
    print('$name', 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'serial', round(d['roofline']['serialized_step_ms'], 2), 'adjoint ms', round(c.get('dcn_col2im(scatter)', {}).get('ms_per_step', 0), 3))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/fh_bench_$name.err').read()[-800:])
PY
}
run new0
run old0 LSNET_LIB_PATH=$OLD
run new1
run old1 LSNET_LIB_PATH=$OLD
