#!/bin/bash
# driver-style validation on one B200: GPU tests, smoke(), bench (own arm incl. cpu_baseline), bench --impl reference
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu -p no:cacheprovider --durations=6 > gpurun_out/pytest_gpu.log 2>&1
grep -E "^E  +|passed|failed|^FAILED|s call|s setup" gpurun_out/pytest_gpu.log | cut -c1-220 | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
cut -c1-1500 gpurun_out/bench_default.json
timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
cut -c1-700 gpurun_out/bench_ref.json
