import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import lsnet_b200 as L
from lsnet_b200.modules import backbone as bb
torch.manual_seed(1)
net = L.build_backbone(dict(type='ResNet', depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1,
                            norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, style='pytorch'))
net.init_weights(None)
for m in net.modules():
    if isinstance(m, torch.nn.BatchNorm2d):
        m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
        m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
net.cuda().train()
x = torch.randn(2, 3, 200, 264, device='cuda')
def run(fuse, only_last=False):
    bb.BWD_FUSE = fuse
    net.zero_grad()
    outs = net(x)
    loss = (outs[-1].float() ** 2).mean() if only_last else sum((o.float() ** 2).mean() for o in outs)
    loss.backward()
    torch.cuda.synchronize()
    return {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
for only_last in (True, False):
    g1 = run(True, only_last); g0 = run(False, only_last)
    print('only_last', only_last)
    for k in g0:
        n0 = float(g0[k].norm())
        if n0 > 0 and k.endswith('conv1.weight') or k.endswith('conv3.weight') or k.endswith('conv2.weight') or 'downsample.0' in k:
            rel = float((g1[k] - g0[k]).norm()) / (n0 + 1e-30)
            print(f'  {k:40s} rel {rel:.4f}')
