import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lsnet_b200.ops.dcn as dcn_mod
from lsnet_b200.data import MODEL_CFG, synthetic_batch, to_device
from lsnet_b200.train import GraphTrainer, Trainer
from lsnet_b200.modules.detector import parse_losses
from lsnet_b200.modules import backbone as bb

batches = [synthetic_batch(s, batch=2, img_hw=(384, 512)) for s in range(3)]
torch.manual_seed(0)
eager = Trainer(MODEL_CFG['bbox_r50'])
sd = {k: v.clone() for k, v in eager.core.state_dict().items()}
graph = GraphTrainer(MODEL_CFG['bbox_r50'], batches[0])
graph.core.load_state_dict(sd)
# 1) forward-only losses on identical weights, eager path on both models
with torch.no_grad():
    l_e = float(parse_losses(eager.core(**to_device(batches[0], 'cuda')))[0])
    l_g = float(parse_losses(graph.core(**to_device(batches[0], 'cuda')))[0])
print('eager-model eager fwd', l_e, 'graph-model eager fwd', l_g)
graph.load_batch(batches[0]); graph.graph.replay(); torch.cuda.synchronize()
print('graph replay loss', float(graph.loss))
graph.graph.replay(); torch.cuda.synchronize()
print('graph replay loss again', float(graph.loss))
graph.load_batch(batches[1]); graph.graph.replay(); torch.cuda.synchronize()
with torch.no_grad():
    l_g1 = float(parse_losses(graph.core(**to_device(batches[1], 'cuda')))[0])
print('batch1: graph replay', float(graph.loss), 'eager fwd', l_g1)
# 2) unfused backbone for comparison
bb._FUSED_OK.update(checked=True, ok=False)
with torch.no_grad():
    print('unfused backbone eager fwd', float(parse_losses(eager.core(**to_device(batches[0], 'cuda')))[0]))
