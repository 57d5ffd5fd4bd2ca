import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lsnet_b200.data import MODEL_CFG, synthetic_batch, to_device
from lsnet_b200.train import GraphTrainer, Trainer, parse_losses
b = synthetic_batch(0, batch=2, img_hw=(384, 512))
torch.manual_seed(0)
eager = Trainer(MODEL_CFG['bbox_r50'])
sd = {k: v.clone() for k, v in eager.core.state_dict().items()}
torch.manual_seed(0)
graph = GraphTrainer(MODEL_CFG['bbox_r50'], b)
graph.core.load_state_dict(sd)
# eager gradients
bd = to_device(b, 'cuda')
eager.model.zero_grad(set_to_none=True)
losses = eager.model(img=bd['img'], img_metas=bd['img_metas'], gt_bboxes=bd['gt_bboxes'], gt_labels=bd['gt_labels'], gt_extremes=bd.get('gt_extremes'))
le, _ = parse_losses(losses)
le.backward()
ge = {k: p.grad.detach().float().clone() for k, p in eager.core.named_parameters() if p.grad is not None}
# graph gradients: replay only (no optimizer step)
graph.load_batch(b)
graph.graph.replay()
torch.cuda.synchronize()
gg = {k: p.grad.detach().float().clone() for k, p in graph.core.named_parameters() if p.grad is not None}
print('loss eager', float(le), 'graph', float(graph.loss))
bad = []
for k in ge:
    if k not in gg:
        print('missing in graph', k); continue
    a, c = ge[k].flatten(), gg[k].flatten()
    na = float(a.norm())
    rel = float((a - c).norm() / (na + 1e-30))
    if na > 1e-9 and rel > 0.05:
        bad.append((rel, k, na, float(c.norm())))
bad.sort(reverse=True)
print('params', len(ge), 'with rel diff > 5%:', len(bad))
for r in bad[:40]:
    print(f'  rel {r[0]:.3f}  {r[1]:60s} |eager| {r[2]:.4e} |graph| {r[3]:.4e}')
# second replay must give the same gradients (flat_g zeroed inside the graph)
graph.graph.replay(); torch.cuda.synchronize()
g2 = {k: p.grad.detach().float().clone() for k, p in graph.core.named_parameters() if p.grad is not None}
worst = max(float((g2[k] - gg[k]).norm() / (gg[k].norm() + 1e-30)) for k in gg)
print('replay-to-replay worst rel diff', worst)
