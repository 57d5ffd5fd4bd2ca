#!/bin/bash
cd /root/repo
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "graph_trainer_matches_eager" 2>&1 | grep -v Warn | grep "assert\|passed\|failed\|Error" | head -5; done
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "groupnorm" 2>&1 | tail -2
