"""One training step of the bench workload under the CUDA profiler API (for ncu --profile-from-start off).
Default: the CUDA-graph step (what bench.py times); `--eager` profiles the per-op eager step."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lsnet_b200.data import MODEL_CFG, synthetic_batch, to_device
from lsnet_b200.train import GraphTrainer, Trainer

eager = '--eager' in sys.argv
cfg_name = sys.argv[sys.argv.index('--config') + 1] if '--config' in sys.argv else 'bbox_r50'
from lsnet_b200.data import TASK_OF
host = [synthetic_batch(s, 0, 4, (800, 1333), pin=True, task=TASK_OF[cfg_name]) for s in range(2)]
if eager:
    tr = Trainer(MODEL_CFG[cfg_name], device='cuda:0')
    batches = [to_device(b, 'cuda:0') for b in host]
else:
    tr = GraphTrainer(MODEL_CFG[cfg_name], host[0], device='cuda:0')
    batches = host
for w in range(3):
    tr.step(batches[w % 2])
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(batches[1])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled 1 step', 'eager' if eager else 'graph')
