import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lsnet_b200.data import MODEL_CFG, synthetic_batch, to_device
from lsnet_b200.train import Trainer, parse_losses
b = to_device(synthetic_batch(0, batch=2, img_hw=(384, 512)), 'cuda')
torch.manual_seed(0)
tr = Trainer(MODEL_CFG['bbox_r50'])
vals = []
with_bwd = '--bwd' in sys.argv
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 30):
    tr.model.zero_grad(set_to_none=True)
    losses = tr.model(img=b['img'], img_metas=b['img_metas'], gt_bboxes=b['gt_bboxes'], gt_labels=b['gt_labels'], gt_extremes=b.get('gt_extremes'))
    loss, _ = parse_losses(losses)
    if with_bwd:
        loss.backward()
    vals.append(round(float(loss), 5))
import collections
print(os.environ.get('TAG', ''), 'distinct losses:', collections.Counter(vals).most_common(6))
