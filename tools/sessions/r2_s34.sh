#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -k "detector_vs_oracle_bbox" -s 2>&1 | grep "seed\|COS\|passed\|failed\|assert\|Error" | head -24
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "graph_trainer_matches_eager" 2>&1 | grep -v Warn | grep "assert\|passed\|failed\|Error" | head -5; done
