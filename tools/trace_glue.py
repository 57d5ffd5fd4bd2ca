"""Which torch (aten) ops still launch kernels inside one eager training step, attributed to the lsnet_b200 source line
that called them (torch.profiler with_stack).  Output: gpurun_out/trace_glue.txt — op, caller, launches, device us."""
import collections
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from lsnet_b200.data import MODEL_CFG, synthetic_batch, to_device
from lsnet_b200.train import Trainer

tr = Trainer(MODEL_CFG['bbox_r50'], device='cuda:0')
batches = [to_device(synthetic_batch(s, 0, 4, (800, 1333)), 'cuda:0') for s in range(2)]
for w in range(3):
    tr.step(batches[w % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True) as prof:
    tr.step(batches[1])
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if not ev.name.startswith('aten::') or not ev.kernels:
        continue
    if any(c.kernels for c in ev.cpu_children if c.name.startswith('aten::')):
        continue                                   # count the innermost aten op that owns the launch
    frame = next((s for s in ev.stack if 'lsnet_b200/' in s), ev.stack[0] if ev.stack else '?')
    frame = frame.split('lsnet_b200/')[-1]
    k = (ev.name, frame)
    agg[k][0] += len(ev.kernels)
    agg[k][1] += sum(kk.duration for kk in ev.kernels)
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
os.makedirs('gpurun_out', exist_ok=True)
with open('gpurun_out/trace_glue.txt', 'w') as f:
    f.write(f'total aten launches {sum(v[0] for v in agg.values())}, device us {sum(v[1] for v in agg.values()):.0f}\n')
    for (name, frame), (n, us) in rows:
        f.write(f'{n:5d} {us:9.1f} us  {name:34s} {frame}\n')
print(open('gpurun_out/trace_glue.txt').read()[:6000])
