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


def _loader_worker(rank, world, port, out):
    """Data side of the N>1 path: every rank builds the SAME dataset and takes its share through build_dataloader
    (rank / world size from the process group, DistributedGroupSampler) -- two epochs through train_epochs with a
    recording trainer standing in for the GPU step."""
    import sys
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import synth_coco as S
    from lsnet_b200 import datasets as D
    from lsnet_b200.registry import DATASETS
    from lsnet_b200.train import train_epochs

    class InjectImage:                      # the fixtures' images are generated, not stored
        def __call__(self, r):
            i = r['img_info']['id'] - 10
            r.update(img=S.image(i), img_shape=S.image(i).shape, ori_shape=S.image(i).shape, img_fields=['img'],
                     filename=r['img_info']['filename'], ori_filename=r['img_info']['filename'])
            return r
    ds = DATASETS.get('CocoDataset')(ann_file=S.coco_dict(False), pipeline=[InjectImage()] + S.pipeline('bbox', True))
    dl = D.build_dataloader(ds, samples_per_gpu=1, workers_per_gpu=0, dist=True, seed=7)
    seen = []

    class Recorder:                         # Trainer protocol: .device, .step(batch)
        device = torch.device('cpu')

        def step(self, batch):
            m = batch['img_metas'][0]
            seen.append((m['ori_filename'], tuple(batch['img'].shape), len(batch['gt_bboxes'][0])))
            return torch.zeros(()), {}
    n = train_epochs(Recorder(), dl, epochs=2)
    out[rank] = (n, seen, type(dl.sampler).__name__)
    dist.destroy_process_group()


def test_distributed_loader_shards_the_dataset():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_loader_worker, args=(2, port, out), nprocs=2, join=True)
    (n0, seen0, kind0), (n1, seen1, kind1) = out[0], out[1]
    assert kind0 == kind1 == 'DistributedGroupSampler'
    # 3 landscape + 3 portrait images over 2 ranks, samples_per_gpu 1: each group is padded to 4 -> 4 steps per rank and epoch
    assert n0 == n1 == 8
    for e in range(2):
        a = [f for f, _, _ in seen0[4 * e:4 * e + 4]]
        b = [f for f, _, _ in seen1[4 * e:4 * e + 4]]
        assert set(a) | set(b) == {f'img_{i}.png' for i in range(6)}       # together the ranks cover the epoch
    assert [f for f, _, _ in seen0[:4]] != [f for f, _, _ in seen0[4:]]     # set_epoch reshuffles
    for _, shape, g in seen0 + seen1:
        assert shape[0] == 1 and shape[1] == 3 and shape[2] % 32 == 0 and shape[3] % 32 == 0 and g >= 1
