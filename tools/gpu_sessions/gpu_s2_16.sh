#!/bin/bash
echo "binned:"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "adjoint" -p no:cacheprovider 2>&1 | grep -E "^E   +Assert|passed|failed"
echo "direct:"; LSNET_COL2IM=direct timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "adjoint" -p no:cacheprovider 2>&1 | grep -E "^E   +Assert|passed|failed"
