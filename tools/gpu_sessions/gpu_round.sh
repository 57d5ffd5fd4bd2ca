#!/bin/bash
# one round of GPU validation: kernel parity groups, model parity, bench
mkdir -p gpurun_out
for grp in "groupnorm" "dcn" "conv2d"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$grp" -rA -p no:cacheprovider > "gpurun_out/pytest_${name}.log" 2>&1
  echo "exit $?" >> "gpurun_out/pytest_${name}.log"
  grep -E "^(FAILED|ERROR)|^E  |passed|failed|exit" "gpurun_out/pytest_${name}.log" | head -20
done
timeout 1200 python -m pytest tests/test_gpu_model.py -q -m gpu -rA -p no:cacheprovider > gpurun_out/pytest_model.log 2>&1
echo "exit $?" >> gpurun_out/pytest_model.log
grep -E "^(FAILED|ERROR)|^E  |passed|failed|exit" gpurun_out/pytest_model.log | head -20
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
for k,v in d['roofline']['classes'].items(): print(' ', k, v)
PY
tail -n 5 gpurun_out/bench.err
