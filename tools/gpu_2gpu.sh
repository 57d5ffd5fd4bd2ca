#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
    print('N2 value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'n_gpus', d['n_gpus'], d['config']['parallelism'])
except Exception as e: print('parse fail', e)
PY
tail -5 gpurun_out/bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 5 --warmup 3 --impl reference > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err; echo "ref n2 exit $?"; cut -c1-300 gpurun_out/bench_n2_ref.json
