#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -k "assign or oracle_bbox or other_tasks" -p no:cacheprovider 2>&1 | grep -E "^E   +|passed|failed"
timeout 600 python tools/trace_step.py > /dev/null 2>&1
sed -n 3p gpurun_out/trace_summary.md; grep "atss" gpurun_out/trace_summary.md
