#!/bin/bash
cd /root/repo
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "Warn\|warn" | grep -B2 -A25 "def test_detector_vs_oracle_bbox\|Error\|^E " | tail -60
