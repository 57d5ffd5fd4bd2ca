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
