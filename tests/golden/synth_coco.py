"""Seeded COCO-format annotation dicts + images for the data-path tests (shared by make_golden_data.py, which runs the
reference pipelines on them, and tests/test_datasets_host.py, which runs ours).  Covers the cases the reference's
annotation parser and contour code branch on: crowd / ignored / degenerate / out-of-image / foreign-class instances,
multi-component polygons, components that are filtered as tiny, polygons with more vertices than the resampling target
(down-sampling branch), and instances whose every component is tiny (box-rectangle fallback)."""
import numpy as np

SIZES = [(64, 88), (90, 60), (72, 72), (48, 96), (80, 56), (66, 90)]        # (height, width); 3 landscape, 3 portrait/square
SCALE = (160, 96)                                                          # the pipelines' img_scale (long, short)
MS_SCALES = [(160, 72), (160, 120)]


def image(i):
    h, w = SIZES[i]
    rng = np.random.RandomState(500 + i)
    base = rng.randint(0, 256, (h // 4 + 2, w // 4 + 2, 3)).astype(np.float32)
    img = np.kron(base, np.ones((4, 4, 1), np.float32))[:h, :w] + rng.randint(-20, 21, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def _ring(rng, cx, cy, rx, ry, n):
    th = np.sort(rng.rand(n)) * 2 * np.pi
    if rng.rand() < 0.5:
        th = th[::-1]                               # both orientations occur in COCO
    r = 0.55 + 0.45 * rng.rand(n)
    return np.stack([cx + rx * r * np.cos(th), cy + ry * r * np.sin(th)], 1).reshape(-1).round(2).tolist()


def _extreme(rng, b):
    x1, y1, x2, y2 = b
    u = rng.rand(4)
    return [x1 + u[0] * (x2 - x1), y1, x1, y1 + u[1] * (y2 - y1), x1 + u[2] * (x2 - x1), y2, x2,
            y1 + u[3] * (y2 - y1), (x1 + x2) / 2, (y1 + y2) / 2]


def _kps(rng, b):
    x1, y1, x2, y2 = b
    v = rng.choice([0, 1, 2], 17, p=[0.25, 0.25, 0.5])
    v[0] = 2
    out = []
    for k in range(17):
        out += [0, 0, 0] if v[k] == 0 else [round(float(x1 + rng.rand() * (x2 - x1)), 2),
                                            round(float(y1 + rng.rand() * (y2 - y1)), 2), int(v[k])]
    return out


def coco_dict(pose=False):
    rng = np.random.RandomState(77 if pose else 42)
    cats = [dict(id=1, name='person'), dict(id=3, name='car'), dict(id=18, name='dog'), dict(id=99, name='not-coco')]
    images, anns = [], []
    aid = 1
    for i, (h, w) in enumerate(SIZES):
        images.append(dict(id=10 + i, file_name=f'img_{i}.png', height=h, width=w))
        for j in range(int(rng.randint(2, 6))):
            bw, bh = rng.uniform(12, w * 0.8), rng.uniform(12, h * 0.8)
            x1, y1 = rng.uniform(0, w - bw), rng.uniform(0, h - bh)
            box = [round(float(x1), 2), round(float(y1), 2), round(float(bw), 2), round(float(bh), 2)]
            cx, cy = x1 + bw / 2, y1 + bh / 2
            kind = (i * 7 + j) % 6
            if kind == 0:       # two components + a tiny one that is filtered
                seg = [_ring(rng, cx - bw / 4, cy, bw / 4, bh / 2, 9), _ring(rng, cx + bw / 4, cy, bw / 5, bh / 3, 14),
                       _ring(rng, cx, cy, 0.8, 0.8, 5)]
            elif kind == 1:     # more vertices than 36 * 10: the down-sampling branch
                seg = [_ring(rng, cx, cy, bw / 2, bh / 2, 400)]
            elif kind == 2:     # every component tiny: falls back to the box rectangle
                seg = [_ring(rng, cx, cy, 0.6, 0.6, 6)]
            else:
                seg = [_ring(rng, cx, cy, bw / 2, bh / 2, int(rng.randint(4, 40)))]
            a = dict(id=aid, image_id=10 + i, category_id=1 if pose else int(rng.choice([1, 3, 18])), bbox=box,
                     area=round(float(bw * bh * 0.6), 2), iscrowd=0, segmentation=seg)
            xyxy = [box[0], box[1], box[0] + box[2], box[1] + box[3]]
            if pose:
                a['keypoints'] = _kps(rng, xyxy)
                a['num_keypoints'] = int(sum(1 for v in a['keypoints'][2::3] if v > 0))
            else:
                a['extreme_points'] = [round(float(v), 2) for v in _extreme(rng, xyxy)]
            anns.append(a)
            aid += 1
        # the rejected kinds (coco.py:141-154), one each on some images
        extra = dict(image_id=10 + i, segmentation=[_ring(rng, w / 2, h / 2, 10, 10, 8)], area=50.0, iscrowd=0,
                     category_id=1, bbox=[5.0, 5.0, 20.0, 20.0])
        extra['keypoints' if pose else 'extreme_points'] = _kps(rng, [5, 5, 25, 25]) if pose else \
            [round(float(v), 2) for v in _extreme(rng, [5, 5, 25, 25])]
        for tweak in (dict(iscrowd=1), dict(ignore=True), dict(bbox=[5.0, 5.0, 0.5, 20.0]), dict(area=0.0),
                      dict(bbox=[w + 5.0, 5.0, 20.0, 20.0]), dict(category_id=99))[i % 3::3]:
            anns.append(dict(extra, id=aid, **tweak))
            aid += 1
    # an image without annotations (filter_empty_gt) and one below min_size
    images.append(dict(id=90, file_name='empty.png', height=64, width=64))
    images.append(dict(id=91, file_name='small.png', height=20, width=64))
    anns.append(dict(extra, id=aid, image_id=91))
    return dict(images=images, annotations=anns, categories=cats)


NORM = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)


def pipeline(task, multiscale=False):
    """The three train pipelines of the LSNet configs (configs/_base_/datasets/coco_lsvr.py:5-14, coco_pose.py:5-14,
    configs/lsnet/lsnet_segm_r50_fpn_1x_coco.py:8-18) at the fixtures' scale; images are injected instead of loaded."""
    load = dict(bbox=dict(type='LoadAnnotations', with_bbox=True, with_extreme=True),
                pose_bbox=dict(type='LoadAnnotations', with_bbox=True, with_keypoint=True),
                segm=dict(type='LoadAnnotations', with_bbox=True, with_mask=True, poly2mask=False, spline_num=10,
                          num_contour_points=36))[task]
    keys = dict(bbox=['img', 'gt_bboxes', 'gt_labels', 'gt_extremes'],
                pose_bbox=['img', 'gt_bboxes', 'gt_labels', 'gt_keypoints'],
                segm=['img', 'gt_bboxes', 'gt_labels', 'gt_masks'])[task]
    resize = dict(type='Resize', img_scale=MS_SCALES, multiscale_mode='range', keep_ratio=True) if multiscale else \
        dict(type='Resize', img_scale=SCALE, keep_ratio=True)
    flip = dict(type='RandomFlip', flip_ratio=0.5)
    if task == 'segm':
        flip['keep_poly_clockwise'] = True
    return [load, resize, flip, dict(type='Normalize', **NORM), dict(type='Pad', size_divisor=32),
            dict(type='DefaultFormatBundle'), dict(type='Collect', keys=keys)]


def eval_pipeline(multi=False):
    """The test pipeline of the LSNet configs (configs/_base_/datasets/coco_lsvr.py:15-29) at the fixtures' scale;
    ``multi``: two scales with flip (the shape of the multi-scale testing configs)."""
    inner = [dict(type='Resize', keep_ratio=True), dict(type='RandomFlip'), dict(type='Normalize', **NORM),
             dict(type='Pad', size_divisor=32), dict(type='ImageToTensor', keys=['img']),
             dict(type='Collect', keys=['img'])]
    return [dict(type='MultiScaleFlipAug', img_scale=MS_SCALES if multi else SCALE, flip=multi, transforms=inner)]


def tta_case(task, seed=0, n_obj=30, num_classes=3):
    """Synthetic per-augmentation detections for the voting tests: ``n_obj`` objects in a 200x300 original image, seen by
    4 augmentations (2 scales x {plain, flipped}) with jitter, random scores and occasional misses — so that clusters of
    1, 2 and 4 members, soft-suppressed leftovers and single-detection classes all occur.  Returns
    ([(bboxes (n,5), vectors (n,V), labels (n,))] as float32/int64 numpy, [meta dict]) in AUGMENTED coordinates."""
    rng = np.random.RandomState(900 + seed)
    H, W = 200, 300
    V = dict(bbox=8, segm=72, pose_bbox=34)[task]
    cx, cy = rng.uniform(30, W - 30, n_obj), rng.uniform(30, H - 30, n_obj)
    bw, bh = rng.uniform(4, 120, n_obj), rng.uniform(4, 90, n_obj)
    boxes = np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
    labels = rng.randint(0, num_classes - 1, n_obj)
    labels[0] = num_classes - 1                                   # a class with exactly one object
    vec = np.stack([boxes[:, [0]] + rng.rand(n_obj, V // 2) * bw[:, None],
                    boxes[:, [1]] + rng.rand(n_obj, V // 2) * bh[:, None]], 2).reshape(n_obj, V)
    dets, metas = [], []
    for a, (scale, flip) in enumerate([(0.8, False), (0.8, True), (1.5, False), (1.5, True)]):
        sf = np.array([scale, scale * 1.01, scale, scale * 1.01], np.float32)
        shape = (int(H * sf[1] + 0.5), int(W * sf[0] + 0.5), 3)
        seen = rng.rand(n_obj) < (1.0 if a == 0 else 0.7)
        seen[0] = a == 2                                          # the lone-class object shows up in one augmentation only
        b = (boxes[seen] + rng.randn(seen.sum(), 4) * 1.5) * sf
        v = (vec[seen] + rng.randn(seen.sum(), V) * 1.0) * np.tile(sf[:2], V // 2)
        if flip:
            b = np.stack([shape[1] - b[:, 2], b[:, 1], shape[1] - b[:, 0], b[:, 3]], 1)
            v = v.copy()
            v[:, 0::2] = shape[1] - v[:, 0::2]                    # (landmark ORDER is whatever the network predicts)
        s = rng.uniform(0.06, 0.95, seen.sum())
        dets.append((np.concatenate([b, s[:, None]], 1).astype(np.float32), v.astype(np.float32),
                     labels[seen].astype(np.int64)))
        metas.append(dict(img_shape=shape, scale_factor=sf, flip=flip, flip_direction='horizontal'))
    return dets, metas


TTA_SCALE_RANGES = [[0, 90], [20, 10000]]
