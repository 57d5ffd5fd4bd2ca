"""The N>1 path on CPU: world_size-2 gloo.  The LSNet kernels need a GPU, so a tiny CPU module stands in for the
detector; what is under test is the host-side data-parallel logic of lsnet_b200.train.Trainer / parse_losses: DDP
gradient averaging == single-process step on the averaged loss, identical parameters on both ranks after the step,
clip + SGD applied after the all-reduce, and the single lazy all-reduce of the logged scalars."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = nn.Linear(6, 5)
        self.b = nn.Linear(5, 3)

    def forward(self, x, y):
        h = torch.relu(self.a(x))
        o = self.b(h)
        return {'loss_cls': [((o - y) ** 2).mean(), (o.abs()).mean() * 0.1], 'loss_bbox_init': (h ** 2).mean()}


CFG = dict(optimizer=dict(lr=0.05, momentum=0.9, weight_decay=1e-4), grad_clip=dict(max_norm=0.5, norm_type=2))


def _data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return dict(x=torch.randn(4, 6, generator=g), y=torch.randn(4, 3, generator=g))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from lsnet_b200.train import Trainer
    tr = Trainer(CFG, device='cpu', distributed=True, model=Tiny())
    tr.iter = 10 ** 6
    logs = None
    for _ in range(3):
        _, logs = tr.step(_data(rank), sync_log=True)
    out[rank] = ({k: v.detach().clone() for k, v in tr.core.state_dict().items()}, dict(logs))
    dist.destroy_process_group()


def test_ddp_step_matches_single_process_average():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    sd0, log0 = out[0]
    sd1, log1 = out[1]
    for k in sd0:
        assert torch.equal(sd0[k], sd1[k]), k                       # replicas stay identical
    assert log0 == log1                                              # one all-reduce, same means on every rank
    # single-process reference: same three steps on the rank-averaged loss
    from lsnet_b200.modules.detector import parse_losses
    from lsnet_b200.train import Trainer
    m = Tiny()
    tr = Trainer(CFG, device='cpu', distributed=False, model=m)
    tr.iter = 10 ** 6
    for _ in range(3):
        tr.optimizer.zero_grad()
        tot = sum(parse_losses(m(**_data(r)))[0] for r in range(2)) / 2
        tot.backward()
        torch.nn.utils.clip_grad_norm_(tr.params, 0.5)
        tr.optimizer.step()
    for k, v in m.state_dict().items():
        assert torch.allclose(v, sd0[k], rtol=1e-5, atol=1e-6), k


def test_warmup_schedule():
    from lsnet_b200.train import warmup_lr
    assert abs(warmup_lr(0.01, 0) - 0.01 * 0.001) < 1e-12       # schedule_1x.py:5-11: warmup_ratio 0.001
    assert warmup_lr(0.01, 500) == 0.01
    assert warmup_lr(0.01, 250) < 0.01
