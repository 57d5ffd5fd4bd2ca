import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, 'tests/golden')
import torch, torch.nn.functional as F
import lsnet_b200 as L
import synth
from lsnet_b200.data import MODEL_CFG
from lsnet_b200.modules import backbone as bb
from oracle import init as oinit
cfg = MODEL_CFG['bbox_r50']
model = L.build_detector(cfg['model'], train_cfg=cfg['train_cfg'])
sd = oinit.make_state_dict('bbox', seed=11)
model.load_state_dict(sd)
model.cuda().train()
d = synth.detector_batch('bbox', 101)
img = d['img'].cuda()
net = model.backbone
outs = {}
for mode in (True, False):
    bb.STEM_OWN = mode
    with torch.no_grad():
        if mode:
            from lsnet_b200 import ops
            wp, shift = ops.pack_stem_weight(net.conv1.weight.detach(), net.bn1.weight.detach(), net.bn1.bias.detach(), net.bn1.running_mean, net.bn1.running_var, net.bn1.eps)
            s = ops.stem_conv(img, wp, shift)
            outs['stem_own'] = s.float()
            outs['pool_own'] = ops.maxpool3x3s2(s).float()
        feats = net(img)
    outs[mode] = [f.float() for f in feats]
w = net.conv1.weight.detach(); bn = net.bn1
sc = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
wf = (w * sc.view(-1, 1, 1, 1)).to(torch.bfloat16).float()
ref = F.relu(F.conv2d(img.to(torch.bfloat16).float(), wf, bn.bias.detach() - bn.running_mean * sc, 2, 3))
print('stem own vs fp32 conv on bf16 operands: max abs', float((outs['stem_own'] - ref).abs().max()), 'ref max', float(ref.abs().max()),
      'mean abs', float((outs['stem_own'] - ref).abs().mean()), 'ref mean', float(ref.abs().mean()))
refp = F.max_pool2d(ref.to(torch.bfloat16).float(), 3, 2, 1)
print('pool: max abs', float((outs['pool_own'] - refp).abs().max()))
for a, b in zip(outs[True], outs[False]):
    print('stage out own-stem vs cudnn-stem: rel max', float((a - b).abs().max() / b.abs().max()), 'rel mean', float((a - b).abs().mean() / b.abs().mean()))
print('bn1 stats', float(bn.running_mean.abs().max()), float(bn.running_var.min()), float(bn.weight.abs().max()), float(w.abs().max()))
