#!/bin/bash
# round 2, last build: whole gpu suite, smoke, the driver's bench line (with the CPU baseline leg), segm config
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/fin_suite.log 2>&1; tail -3 gpurun_out/fin_suite.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/fin_bench_default.json 2> gpurun_out/fin_bench_default.err; tail -c 1500 gpurun_out/fin_bench_default.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/fin_bench_20.json 2> gpurun_out/fin_bench_20.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config segm_r50 > gpurun_out/fin_bench_segm.json 2> gpurun_out/fin_bench_segm.err
python - <<'PY'
import json
for n in ('default', '20', 'segm'):
    try:
        d = json.loads(open(f'gpurun_out/fin_bench_{n}.json').read().strip().splitlines()[-1])
        print(n, 'img/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'h2d', d['e2e']['h2d_bytes_per_step'],
              'frac', round(d['roofline']['frac'], 3), 'launches', d['gpu_launches'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(n, 'FAILED', e); print(open(f'gpurun_out/fin_bench_{n}.err').read()[-800:])
PY
