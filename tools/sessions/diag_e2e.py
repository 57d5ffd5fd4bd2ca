import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from lsnet_b200.data import MODEL_CFG, synthetic_batch
from lsnet_b200.train import GraphTrainer
host = [synthetic_batch(s, 0, 4, (800, 1333), pin=True) for s in range(4)]
tr = GraphTrainer(MODEL_CFG['bbox_r50'], host[0], device='cuda:0')
for w in range(4):
    tr.step(host[w % 4], next_batch=host[(w + 1) % 4])[0].item()
def run(prefetch, n=24):
    ts = []
    for s in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if prefetch:
            loss, _ = tr.step(host[s % 4], next_batch=host[(s + 1) % 4])
        else:
            loss, _ = tr.step(host[s % 4])
        t1 = time.perf_counter()
        loss.item()
        t2 = time.perf_counter()
        ts.append((round(1e3 * (t1 - t0), 2), round(1e3 * (t2 - t0), 2)))
    return ts
for rep in range(3):
    a = run(True)
    print('prefetch  (enqueue ms, total ms):', a)
b = run(False)
print('no prefetch:', b)
# H2D bandwidth alone
img = host[0]['img']; dst = torch.empty_like(img, device='cuda')
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(img, non_blocking=True); torch.cuda.synchronize()
    print('H2D 51.6 MB ms', round(1e3 * (time.perf_counter() - t0), 2))
