#!/bin/bash
mkdir -p gpurun_out
for grp in "dcn"; do
  timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$grp" -rA -p no:cacheprovider > "gpurun_out/pytest_${grp}.log" 2>&1
  echo "exit $?" >> "gpurun_out/pytest_${grp}.log"
  grep -E "^(FAILED|ERROR)|^E  |passed|failed|exit" "gpurun_out/pytest_${grp}.log" | head -20
done
timeout 1200 python -m pytest tests/test_gpu_model.py -q -m gpu -rA -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1
echo "exit $?" >> gpurun_out/pytest_model.log
grep -E "^(FAILED|ERROR)|^E  |passed|failed|exit" gpurun_out/pytest_model.log | head -20
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
for k,v in d['roofline']['classes'].items(): print(' ', k, v)
print('cpu', d['cpu_baseline'])
PY
tail -n 3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-700 gpurun_out/bench_ref.json
