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


def _derive(base, **over):
    import copy
    c = copy.deepcopy(base)
    for k, v in over.items():
        tgt = c
        *path, last = k.split('.')
        for q in path:
            tgt = tgt[q]
        tgt[last] = v
    return c


_X101_DCN = dict(type='ResNeXt', depth=101, groups=64, base_width=4, num_stages=4, out_indices=(0, 1, 2, 3),
                 frozen_stages=1, norm_cfg=dict(type='BN', requires_grad=True),
                 dcn=dict(type='DCNv2', deformable_groups=1, fallback_on_stride=False),
                 stage_with_dcn=(False, True, True, True), norm_eval=True, with_cp=True, style='pytorch')
_GN = dict(type='GN', num_groups=32, requires_grad=True)
_HEAD = MODEL_CFG['bbox_r50']['model']['bbox_head']
# BASELINE.json configs[2]: configs/lsnet/lsnet_bbox_x101_fpn_dconv_c3-c5_mstrain_2x_coco.py.py (multi-scale 480-960)
MODEL_CFG['bbox_x101dcn_ms'] = _derive(MODEL_CFG['bbox_r50'], **{'model.backbone': _X101_DCN})
# configs[3]: configs/lsnet/lsnet_segm_r50_fpn_1x_coco.py (36 contour landmarks)
MODEL_CFG['segm_r50'] = _derive(MODEL_CFG['bbox_r50'], **{'model.bbox_head': dict(
    type='LSHead', task='segm', num_vectors=36, num_classes=80, in_channels=256, feat_channels=256,
    point_feat_channels=256, stacked_convs=3, num_kernel_points=9, gradient_mul=0.1, point_strides=[8, 16, 32, 64, 128],
    point_base_scale=4, norm_cfg=_GN, conv_module_type='dcn', loss_cls=_HEAD['loss_cls'],
    loss_segm_init=dict(type='CrossIOULoss', loss_weight=1.0, loss_type='polygon', stride=9),
    loss_segm_refine=dict(type='CrossIOULoss', loss_weight=2.0, loss_type='polygon', stride=9))})
# configs[4]: configs/lsnet/lsnet_pose_bbox_x101_fpn_dconv_c3-c5_mstrain_2x_coco.py (17 keypoints, person class only)
MODEL_CFG['pose_x101dcn'] = _derive(MODEL_CFG['bbox_r50'], **{'model.backbone': _X101_DCN, 'model.bbox_head': dict(
    type='LSHead', task='pose_bbox', num_vectors=17, num_classes=1, in_channels=256, feat_channels=256,
    point_feat_channels=256, stacked_convs=3, num_kernel_points=9, gradient_mul=0.1, point_strides=[8, 16, 32, 64, 128],
    point_base_scale=4, norm_cfg=_GN, conv_module_type='dcn', loss_cls=_HEAD['loss_cls'],
    loss_bbox_init=dict(type='CrossIOULoss', loss_weight=0.1, loss_type='bbox'),
    loss_bbox_refine=dict(type='CrossIOULoss', loss_weight=0.2, loss_type='bbox'),
    loss_pose_init=dict(type='CrossIOULoss', loss_weight=1.0, loss_type='keypoint'),
    loss_pose_refine=dict(type='CrossIOULoss', loss_weight=2.0, loss_type='keypoint'))})
TASK_OF = {'bbox_r50': 'bbox', 'bbox_x101dcn_ms': 'bbox', 'segm_r50': 'segm', 'pose_x101dcn': 'pose_bbox'}


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


def _contours(rng, b, n=36):
    """n clockwise contour points per box starting at the top (LoadAnnotations(num_contour_points=36) conventions,
    mmdet/datasets/pipelines/loading.py:408-441), as the (G, 2n+2) table LSHead.process_polygons produces (extent
    centre appended) plus the extent boxes that replace gt_bboxes for task 'segm' (lsnet_head.py:1717-1756)."""
    th = -np.pi / 2 + 2 * np.pi * np.arange(n) / n
    x1, y1, x2, y2 = b.T
    cx, cy = (x1 + x2) / 2, (y1 + y2) / 2
    r = 0.6 + 0.4 * rng.rand(len(b), n)
    P = np.stack([cx[:, None] + r * ((x2 - x1) / 2)[:, None] * np.cos(th)[None],
                  cy[:, None] + r * ((y2 - y1) / 2)[:, None] * np.sin(th)[None]], 2).astype(np.float32)   # (G, n, 2)
    lo, hi = P.min(1), P.max(1)
    ct = (lo + hi) / 2
    table = np.concatenate([P.reshape(len(b), -1), ct], 1).astype(np.float32)
    return table, np.concatenate([lo, hi], 1).astype(np.float32)


def _keypoints(rng, b, n=17):
    """(G, 3n) COCO keypoints [x, y, v] inside the box, v in {0, 1, 2} (coordinates zero when v == 0)."""
    x1, y1, x2, y2 = b.T
    x = x1[:, None] + rng.rand(len(b), n) * (x2 - x1)[:, None]
    y = y1[:, None] + rng.rand(len(b), n) * (y2 - y1)[:, None]
    v = rng.choice([0, 1, 2], size=(len(b), n), p=[0.2, 0.3, 0.5]).astype(np.float32)
    v[:, 0] = 2                       # at least one visible keypoint per instance
    return np.stack([x * (v > 0), y * (v > 0), v], 2).reshape(len(b), 3 * n).astype(np.float32)


def ms_size(rng, short_range=(480, 960), long_max=1333, aspect=4 / 3):
    """Resize(img_scale=[(1333, 480), (1333, 960)], multiscale_mode='range', keep_ratio=True) of a 4:3 COCO image
    (configs/lsnet/lsnet_bbox_r50_fpn_mstrain_2x_coco.py:10-16, mmdet/datasets/pipelines/transforms.py random_sample):
    the short side is drawn uniformly, the long side follows the aspect ratio and is capped at 1333."""
    short = int(rng.randint(short_range[0], short_range[1] + 1))
    scale = min(short / 1.0, long_max / aspect)
    return int(round(scale)), int(round(scale * aspect))


#: img_norm_cfg of every LSNet config (configs/_base_/datasets/coco_lsvr.py:3-4)
IMG_NORM_CFG = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)


def synthetic_batch(step, rank=0, batch=4, img_hw=(800, 1333), divisor=32, task='bbox', pin=False, multiscale=None,
                    canvas_multiple=None, u8=False):
    """One per-GPU batch: dict(img [B,3,Hp,Wp] fp32, img_metas, gt_bboxes, gt_labels, + task ground truth:
    'bbox' gt_extremes (G,10); 'segm' gt_masks = (G,74) contour tables (LSHead.process_polygons accepts them) and
    gt_bboxes = contour extents; 'pose_bbox' gt_keypoints (G,51), labels 0).
    ``multiscale=(lo, hi)``: every image gets its own size (short side in [lo, hi], collate pads to the largest,
    mmcv/parallel/collate.py:39-60); ``canvas_multiple`` additionally rounds the canvas up (shape buckets of the CUDA-graph
    cache).  ``u8``: the image is the uint8 [B, Hp, Wp, 3] byte batch of the device-prep pipeline
    (datasets.loader.collate) with ``img_hw`` / ``img_norm_cfg``; the trainer normalises it on the GPU."""
    rng = np.random.RandomState(1234 + 8 * step + rank)
    sizes = [ms_size(rng, multiscale) if multiscale else tuple(img_hw) for _ in range(batch)]
    pads = [((h + divisor - 1) // divisor * divisor, (w + divisor - 1) // divisor * divisor) for h, w in sizes]
    Hp, Wp = max(p[0] for p in pads), max(p[1] for p in pads)
    if canvas_multiple:
        Hp = (Hp + canvas_multiple - 1) // canvas_multiple * canvas_multiple
        Wp = (Wp + canvas_multiple - 1) // canvas_multiple * canvas_multiple
    img = np.zeros((batch, Hp, Wp, 3), np.uint8) if u8 else np.zeros((batch, 3, Hp, Wp), np.float32)
    out = dict(gt_bboxes=[], gt_labels=[])
    if task == 'bbox':
        out['gt_extremes'] = []
    elif task == 'segm':
        out['gt_masks'] = []
    elif task == 'pose_bbox':
        out['gt_keypoints'] = []
    else:
        raise ValueError(task)
    for i, (H, W) in enumerate(sizes):
        if u8:
            img[i, :H, :W] = rng.randint(0, 256, (H, W, 3), dtype=np.uint8)
        else:
            img[i, :, :H, :W] = rng.rand(3, H, W).astype(np.float32)
        b = _boxes(rng, rng.randint(1, 16), H, W)
        labels = rng.randint(0, 80, len(b)).astype(np.int64)
        if task == 'bbox':
            out['gt_extremes'].append(torch.from_numpy(_extremes(rng, b)))
        elif task == 'segm':
            table, b = _contours(rng, b)
            out['gt_masks'].append(torch.from_numpy(table))
        else:
            out['gt_keypoints'].append(torch.from_numpy(_keypoints(rng, b)))
            labels = np.zeros(len(b), np.int64)
        out['gt_bboxes'].append(torch.from_numpy(b))
        out['gt_labels'].append(torch.from_numpy(labels))
    img = torch.from_numpy(img)
    if pin:
        img = img.pin_memory()
    out['img'] = img
    out['img_metas'] = [dict(img_shape=(H, W, 3), pad_shape=(ph, pw, 3), scale_factor=1.0, flip=False)
                        for (H, W), (ph, pw) in zip(sizes, pads)]
    if u8:
        out['img_hw'] = torch.tensor([[H, W] for H, W in sizes], dtype=torch.int32)
        out['img_norm_cfg'] = IMG_NORM_CFG
        if pin:
            out['img_hw'] = out['img_hw'].pin_memory()
    return out


def to_device(batch, device, non_blocking=True):
    out = dict(batch)
    out['img'] = batch['img'].to(device, non_blocking=non_blocking)
    return out   # GT lists stay on the host: LSHead.loss packs them into one padded tensor per field and uploads that
