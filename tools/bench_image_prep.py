"""lsnet_image_prep_u8 at the bench canvas (B4 800x1344): CUDA-event time with L2 flushed, algorithmic GB/s; also the
harness `ncu -k regex:image_prep` attaches to."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lsnet_b200.data import IMG_NORM_CFG  # noqa: E402
from lsnet_b200.datasets import DevicePrep  # noqa: E402

sets = []
for _ in range(4):           # 4 x 64.5 MB of operands: every launch misses the 126 MB L2
    sets.append((torch.randint(0, 256, (4, 800, 1344, 3), dtype=torch.uint8, device='cuda'),
                 torch.empty((4, 3, 800, 1344), device='cuda').contiguous(memory_format=torch.channels_last)))
hw = torch.tensor([[800, 1333]] * 4, dtype=torch.int32, device='cuda')
p = DevicePrep('cuda')
for u8, out in sets:
    p.run(u8, hw, IMG_NORM_CFG, out=out)
g = torch.cuda.CUDAGraph()       # back-to-back launches (a lone launch measures the host's launch gap, not the kernel)
with torch.cuda.graph(g):
    for _ in range(10):
        for u8, out in sets:
            p.run(u8, hw, IMG_NORM_CFG, out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g.replay()
torch.cuda.synchronize()
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 40
b = sets[0][0].numel() + sets[0][1].numel() * 4
print(f'image_prep_u8 B4 800x1344: {t * 1e3:.1f} us per launch (40 back-to-back, operands rotate through 258 MB), '
      f'{b / 1e6:.1f} MB algorithmic, {b / t / 1e6:.0f} GB/s')
