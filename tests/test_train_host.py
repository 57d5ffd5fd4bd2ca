"""Host-side step logic (no GPU): gradient clip resolved from the reference's ``optimizer_config``, the step / warm-up
learning-rate policy of schedule_1x, and the key-box ground truth of task 'pose_kbox'."""
import os

import pytest
import torch

import lsnet_b200 as L
from lsnet_b200.train import LrSchedule, Trainer, grad_clip_of

REF_CFG = '/root/reference/code/configs/lsnet/lsnet_bbox_r50_fpn_1x_coco.py'


def test_grad_clip_resolution():
    assert grad_clip_of(dict(optimizer_config=dict(grad_clip=dict(max_norm=35, norm_type=2)))) == 35
    assert grad_clip_of(dict(grad_clip=dict(max_norm=10))) == 10
    assert grad_clip_of(dict(optimizer_config=dict(grad_clip=None))) is None
    assert grad_clip_of(dict()) is None
    with pytest.raises(ValueError):
        grad_clip_of(dict(optimizer_config=dict(grad_clip=dict(max_norm=35, norm_type=1))))


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason='reference tree not present')
def test_trainer_from_reference_config_clips_at_35():
    """configs/lsnet/lsnet_bbox_r50_fpn_1x_coco.py:65 puts the clip under optimizer_config."""
    c = L.Config.fromfile(REF_CFG)
    c.model.pretrained = None
    tr = Trainer(c, device='cpu')
    assert tr.max_norm == 35
    assert tr.base_lr == 0.01 and tr.optimizer.defaults['momentum'] == 0.9 and tr.optimizer.defaults['weight_decay'] == 1e-4
    # lr_config of _base_/schedules/schedule_1x.py: linear warm-up over 500 iterations from 0.001, steps at epochs 8, 11
    assert tr.lr_at.warmup_iters == 500 and tr.lr_at.steps == [8, 11]


def test_lr_schedule_matches_mmcv_formulas():
    """mmcv/runner/hooks/lr_updater.py:62-80 (warm-up) and :144-172 (step): regular lr = base * gamma^(#steps passed);
    linear warm-up lr = regular * (1 - (1 - it/warmup_iters) * (1 - ratio))."""
    s = LrSchedule(0.01, dict(policy='step', warmup='linear', warmup_iters=500, warmup_ratio=0.001, step=[8, 11]),
                   iters_per_epoch=100)
    assert abs(s(0) - 0.01 * 0.001) < 1e-12
    assert abs(s(250) - 0.01 * (1 - 0.5 * 0.999)) < 1e-12
    assert s(500) == 0.01 and s(799) == 0.01
    assert abs(s(800) - 0.001) < 1e-12 and abs(s(1099) - 0.001) < 1e-12
    assert abs(s(1100) - 0.0001) < 1e-12
    # without an epoch length the lr stays at its base value after the warm-up
    assert LrSchedule(0.01, dict(policy='step', step=[8, 11]))(10 ** 6) == 0.01
    with pytest.raises(ValueError):
        LrSchedule(0.01, dict(policy='cyclic'))


def test_keybox_ground_truth():
    """LSHead.process_keypoints_with_kbox (lsnet_head.py:1787-1828) on literal keypoints: the box is the extent of the
    VISIBLE keypoints, the appended centre its midpoint, hidden keypoints keep their coordinates."""
    from lsnet_b200.modules.head import LSHead
    k = torch.tensor([[10., 20., 2., 50., 5., 0., 30., 40., 1.],        # middle keypoint hidden
                      [1., 2., 2., 3., 4., 2., 5., 6., 2.]])
    k0 = k.clone()
    kps, boxes, vs = LSHead.process_keypoints_with_kbox([k])
    assert torch.equal(k, k0)                                          # input untouched
    assert torch.equal(boxes[0], torch.tensor([[10., 20., 30., 40.], [1., 2., 5., 6.]]))
    assert torch.equal(vs[0], torch.tensor([[2., 0., 1.], [2., 2., 2.]]))
    assert torch.equal(kps[0], torch.tensor([[10., 20., 50., 5., 30., 40., 20., 30.], [1., 2., 3., 4., 5., 6., 3., 4.]]))


def test_checkpoint_resume_roundtrip(tmp_path):
    """save_checkpoint / resume (mmcv checkpoint layout): a trainer resumed from a file continues exactly like the one
    that wrote it -- weights, momentum, iteration (learning-rate schedule) -- and the file has the reference's keys."""
    import torch
    import torch.nn as nn
    from lsnet_b200.train import Trainer, resume, save_checkpoint

    class Tiny(nn.Module):
        def __init__(self):
            super().__init__()
            self.a = nn.Linear(6, 5)
            self.frozen = nn.Linear(5, 5)
            self.b = nn.Linear(5, 3)
            for p in self.frozen.parameters():
                p.requires_grad = False

        def forward(self, x, y):
            return {'loss_cls': ((self.b(self.frozen(torch.relu(self.a(x)))) - y) ** 2).mean()}
    cfg = dict(optimizer=dict(lr=0.05, momentum=0.9, weight_decay=1e-4), grad_clip=dict(max_norm=0.5, norm_type=2),
               lr_config=dict(policy='step', warmup='linear', warmup_iters=5, warmup_ratio=0.1, step=[1]))
    g = torch.Generator().manual_seed(0)
    data = [dict(x=torch.randn(4, 6, generator=g), y=torch.randn(4, 3, generator=g)) for _ in range(6)]
    torch.manual_seed(1)
    t1 = Trainer(cfg, device='cpu', model=Tiny(), iters_per_epoch=4)
    for d in data[:3]:
        t1.step(d)
    meta = save_checkpoint(t1, str(tmp_path / 'e.pth'), epoch=0, meta=dict(note='x'))
    assert meta == dict(note='x', epoch=0, iter=3)
    ck = torch.load(str(tmp_path / 'e.pth'), weights_only=False)
    assert set(ck) == {'meta', 'state_dict', 'optimizer'} and set(ck['state_dict']) == set(t1.core.state_dict())
    # SGD state over ALL parameters (as the reference's optimizer): the frozen layer's two have no buffer
    assert sorted(ck['optimizer']['state']) == [0, 1, 4, 5] and ck['optimizer']['param_groups'][0]['params'] == list(range(6))
    torch.manual_seed(2)
    t2 = Trainer(cfg, device='cpu', model=Tiny(), iters_per_epoch=4)
    assert resume(t2, str(tmp_path / 'e.pth'))['iter'] == 3 and t2.iter == 3
    for d in data[3:]:
        l1, _ = t1.step(d)
        l2, _ = t2.step(d)
        assert torch.equal(l1, l2)
    for (k, a), (_, b) in zip(t1.core.state_dict().items(), t2.core.state_dict().items()):
        assert torch.equal(a, b), k
