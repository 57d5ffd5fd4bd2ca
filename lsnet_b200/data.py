"""Synthetic COCO-shaped training batches (recipe of the reference's ``_demo_mm_inputs``,
tests/test_forward.py:278-344): uniform random pixels, 1-15 random boxes per image, extreme points / contours /
keypoints derived from the boxes.  ``RandomState(1234 + 8*step + rank)``.  Host tensors (pinned when asked)."""
import numpy as np
import torch

MODEL_CFG = {
    # configs/lsnet/lsnet_bbox_r50_fpn_1x_coco.py, restated as data (the reference file itself loads unmodified
    # through lsnet_b200.Config when the reference tree is present)
    'bbox_r50': dict(
        model=dict(
            type='LSDetector', pretrained=None,
            backbone=dict(type='ResNet', depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1,
                          norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, style='pytorch'),
            neck=dict(type='FPN', in_channels=[256, 512, 1024, 2048], out_channels=256, start_level=1,
                      add_extra_convs='on_input', num_outs=5, norm_cfg=dict(type='GN', num_groups=32, requires_grad=True)),
            bbox_head=dict(type='LSHead', task='bbox', num_vectors=4, num_classes=80, in_channels=256, feat_channels=256,
                           point_feat_channels=256, stacked_convs=3, num_kernel_points=9, gradient_mul=0.1,
                           point_strides=[8, 16, 32, 64, 128], point_base_scale=4,
                           norm_cfg=dict(type='GN', num_groups=32, requires_grad=True), conv_module_type='dcn',
                           loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
                           loss_bbox_init=dict(type='CrossIOULoss', loss_weight=1.0),
                           loss_bbox_refine=dict(type='CrossIOULoss', loss_weight=2.0))),
        train_cfg=dict(init=dict(assigner=dict(type='CentroidAssigner', scale=4, pos_num=1, iou_type='center'),
                                 allowed_border=-1, pos_weight=-1, debug=False),
                       refine=dict(assigner=dict(type='ATSSAssigner', topk=9), allowed_border=-1, pos_weight=-1,
                                   debug=False)),
        optimizer=dict(type='SGD', lr=0.01, momentum=0.9, weight_decay=0.0001),
        grad_clip=dict(max_norm=35, norm_type=2)),
}


def _boxes(rng, G, H, W, min_side=8):
    cx, cy, bw, bh = rng.rand(G, 4).T
    x1 = (cx * W - W * bw / 2).clip(0, W); x2 = (cx * W + W * bw / 2).clip(0, W)
    y1 = (cy * H - H * bh / 2).clip(0, H); y2 = (cy * H + H * bh / 2).clip(0, H)
    b = np.stack([x1, y1, x2, y2], 1).astype(np.float32)
    keep = ((b[:, 2] - b[:, 0]) > min_side) & ((b[:, 3] - b[:, 1]) > min_side)
    if not keep.any():
        return np.array([[W * 0.25, H * 0.25, W * 0.75, H * 0.75]], np.float32)
    return b[keep]


def _extremes(rng, b):
    G = len(b)
    u = rng.rand(G, 4).astype(np.float32)
    x1, y1, x2, y2 = b.T
    return np.stack([x1 + u[:, 0] * (x2 - x1), y1, x1, y1 + u[:, 1] * (y2 - y1), x1 + u[:, 2] * (x2 - x1), y2, x2,
                     y1 + u[:, 3] * (y2 - y1), (x1 + x2) / 2, (y1 + y2) / 2], 1).astype(np.float32)


def synthetic_batch(step, rank=0, batch=4, img_hw=(800, 1333), divisor=32, task='bbox', pin=False):
    """One per-GPU batch: dict(img [B,3,Hp,Wp] fp32, img_metas, gt_bboxes, gt_labels, gt_extremes)."""
    rng = np.random.RandomState(1234 + 8 * step + rank)
    H, W = img_hw
    Hp, Wp = (H + divisor - 1) // divisor * divisor, (W + divisor - 1) // divisor * divisor
    img = np.zeros((batch, 3, Hp, Wp), np.float32)
    img[:, :, :H, :W] = rng.rand(batch, 3, H, W).astype(np.float32)
    gt_b, gt_l, gt_e = [], [], []
    for _ in range(batch):
        b = _boxes(rng, rng.randint(1, 16), H, W)
        gt_b.append(torch.from_numpy(b))
        gt_l.append(torch.from_numpy(rng.randint(0, 80, len(b)).astype(np.int64)))
        gt_e.append(torch.from_numpy(_extremes(rng, b)))
    img = torch.from_numpy(img)
    if pin:
        img = img.pin_memory()
    metas = [dict(img_shape=(H, W, 3), pad_shape=(Hp, Wp, 3), scale_factor=1.0, flip=False) for _ in range(batch)]
    assert task == 'bbox'
    return dict(img=img, img_metas=metas, gt_bboxes=gt_b, gt_labels=gt_l, gt_extremes=gt_e)


def to_device(batch, device, non_blocking=True):
    out = dict(batch)
    out['img'] = batch['img'].to(device, non_blocking=non_blocking)
    return out   # GT lists stay on the host: LSHead.loss packs them into one padded tensor per field and uploads that
