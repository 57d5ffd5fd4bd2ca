import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from lsnet_b200.data import MODEL_CFG, synthetic_batch
from lsnet_b200.train import GraphTrainer
host = [synthetic_batch(s, 0, 4, (800, 1333), pin=True) for s in range(4)]
tr = GraphTrainer(MODEL_CFG['bbox_r50'], host[0], device='cuda:0')
for w in range(4):
    tr.step(host[w % 4], next_batch=host[(w + 1) % 4])[0].item()
def run(n=40):
    ts = []
    for s in range(n):
        t0 = time.perf_counter()
        loss, _ = tr.step(host[s % 4], next_batch=host[(s + 1) % 4])
        loss.item()
        ts.append(round(1e3 * (time.perf_counter() - t0), 1))
    return ts
print('no sampler ', run())
c = bench.ClockSampler(0); c.start()
print('sampler    ', run())
print('sampler    ', run())
print(c.stop())
import pynvml as n
n.nvmlInit(); h = n.nvmlDeviceGetHandleByIndex(0)
for f, name in ((lambda: n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), 'clock'), (lambda: n.nvmlDeviceGetCurrentClocksEventReasons(h), 'reasons')):
    t0 = time.perf_counter()
    for _ in range(20): f()
    print(name, 'ms per query', round(1e3 * (time.perf_counter() - t0) / 20, 3))
