"""One training step of the bench workload under the CUDA profiler API (for ncu --profile-from-start off)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lsnet_b200.data import MODEL_CFG, synthetic_batch, to_device
from lsnet_b200.train import Trainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
tr = Trainer(MODEL_CFG['bbox_r50'], device='cuda:0')
batches = [to_device(synthetic_batch(s, 0, 4, (800, 1333)), 'cuda:0') for s in range(2)]
for w in range(3):
    tr.step(batches[w % 2])
torch.cuda.synchronize()
torch.cuda.profiler.start()
for s in range(steps):
    tr.step(batches[s % 2])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled', steps, 'step(s)')
