#!/bin/bash
run() { env "$@" timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'ms/step', round(d['ms_per_step'],2), 'img/s', round(d['value'],1))"; }
run A=base
run LSNET_IM2COL_U=2
run LSNET_BIN_VARIANT=34
run LSNET_BIN_VARIANT=38
run LSNET_BIN_VARIANT=28
run LSNET_WGRAD_MT2=1
run LSNET_GEMM_ADAPT_BN=1
run LSNET_OVERLAP_WGRAD=1
run LSNET_TOWER_STREAMS=0
run A=base2
