#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/fin_bench_n2.json 2> gpurun_out/fin_bench_n2.err; echo "bench n2 exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/fin_bench_n2.json').read().strip().splitlines()[-1])
    print('N2 value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['e2e'].get('step_wall_ms'), 'n_gpus', d['n_gpus'], d['config']['parallelism'])
except Exception as e: print('parse fail', e)
PY
tail -4 gpurun_out/fin_bench_n2.err

