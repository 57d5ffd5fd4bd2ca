#!/bin/bash
cd /root/repo
TAG=default timeout 300 python tools/sessions/diag_fwd_det.py 40 2>&1 | grep distinct
TAG=default+bwd timeout 300 python tools/sessions/diag_fwd_det.py 40 --bwd 2>&1 | grep distinct
TAG=stem_cudnn LSNET_STEM_OWN=0 timeout 300 python tools/sessions/diag_fwd_det.py 40 --bwd 2>&1 | grep distinct
TAG=nofoldside LSNET_FOLD_SIDE=0 timeout 300 python tools/sessions/diag_fwd_det.py 40 --bwd 2>&1 | grep distinct
TAG=trunk_cudnn LSNET_TRUNK=cudnn timeout 300 python tools/sessions/diag_fwd_det.py 40 --bwd 2>&1 | grep distinct
TAG=nostreams LSNET_LEVEL_STREAMS=0 LSNET_TOWER_STREAMS=0 timeout 300 python tools/sessions/diag_fwd_det.py 40 --bwd 2>&1 | grep distinct
