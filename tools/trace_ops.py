"""aten-op level view of ONE eager training step (torch.profiler, shapes recorded): which torch element-wise / copy ops
surround the library kernels, grouped by input shape."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from lsnet_b200.data import MODEL_CFG, synthetic_batch, to_device
from lsnet_b200.train import Trainer

tr = Trainer(MODEL_CFG['bbox_r50'], device='cuda:0')
batches = [to_device(synthetic_batch(s, 0, 4, (800, 1333)), 'cuda:0') for s in range(2)]
for w in range(3):
    tr.step(batches[w % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    tr.step(batches[1])
    torch.cuda.synchronize()
os.makedirs('gpurun_out', exist_ok=True)
with open('gpurun_out/trace_ops.txt', 'w') as f:
    f.write(prof.key_averages(group_by_input_shape=True).table(sort_by='self_cuda_time_total', row_limit=90,
                                                                max_name_column_width=40, max_shapes_column_width=70))
    f.write('\n\n')
    f.write(prof.key_averages().table(sort_by='self_cuda_time_total', row_limit=60, max_name_column_width=50))
print(open('gpurun_out/trace_ops.txt').read()[-9000:])
