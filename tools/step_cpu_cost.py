"""Host-side cost of one GraphTrainer.step with a per-step sync (the e2e mode): load_batch / replay / optimizer."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lsnet_b200.data import MODEL_CFG, synthetic_batch
from lsnet_b200.train import GraphTrainer
host = [synthetic_batch(s, 0, 4, (800, 1333), pin=True) for s in range(4)]
tr = GraphTrainer(MODEL_CFG['bbox_r50'], host[0], device='cuda:0')
for w in range(5):
    tr.step(host[w % 4]); torch.cuda.synchronize()
acc = dict(load=0.0, replay_call=0.0, replay_gpu=0.0, opt_call=0.0, total=0.0)
N = 10
for s in range(N):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tr.load_batch(host[s % 4]); t1 = time.perf_counter()
    torch.cuda.synchronize(); t1b = time.perf_counter()
    tr.graph.replay(); t2 = time.perf_counter()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    tr.step(None) if False else None
    acc['load'] += t1 - t0; acc['replay_call'] += t2 - t1b; acc['replay_gpu'] += t3 - t1b; acc['total'] += t3 - t0
    acc['opt_call'] += 0
print({k: round(1e3 * v / N, 2) for k, v in acc.items()}, 'ms; load sync wait', )
# whole step with item() like bench e2e
torch.cuda.synchronize(); t0 = time.perf_counter()
for s in range(N):
    loss, _ = tr.step(host[s % 4]); loss.item()
print('e2e-style step', round(1e3 * (time.perf_counter() - t0) / N, 2), 'ms')
torch.cuda.synchronize(); t0 = time.perf_counter()
for s in range(N):
    loss, _ = tr.step(host[s % 4])
torch.cuda.synchronize()
print('async step', round(1e3 * (time.perf_counter() - t0) / N, 2), 'ms')
