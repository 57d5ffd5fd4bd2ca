"""Seeded synthetic inputs shared by tests/golden/make_golden.py (which runs the REFERENCE on them) and the tests
(which run the oracle / the CUDA path on the identical inputs).  Recipe follows the reference's own
``_demo_mm_inputs`` (tests/test_forward.py:278-344) and SURVEY.md §8(d)."""
import numpy as np
import torch


def boxes(rng, G, H, W, min_side=8):
    cx, cy, bw, bh = rng.rand(G, 4).T
    x1 = (cx * W - W * bw / 2).clip(0, W); x2 = (cx * W + W * bw / 2).clip(0, W)
    y1 = (cy * H - H * bh / 2).clip(0, H); y2 = (cy * H + H * bh / 2).clip(0, H)
    b = np.stack([x1, y1, x2, y2], 1).astype(np.float32)
    keep = ((b[:, 2] - b[:, 0]) > min_side) & ((b[:, 3] - b[:, 1]) > min_side)
    if not keep.any():
        b = np.array([[W * 0.25, H * 0.25, W * 0.75, H * 0.75]], np.float32)
        keep = np.array([True])
    return torch.from_numpy(b[keep])


def extremes(rng, b):
    """(G,10) [top, left, bottom, right, centre] as (x,y) (tools/gen_coco_lsvr.py:79,104-107)."""
    G = len(b)
    u = rng.rand(G, 4).astype(np.float32)
    x1, y1, x2, y2 = b.numpy().T
    e = np.stack([x1 + u[:, 0] * (x2 - x1), y1, x1, y1 + u[:, 1] * (y2 - y1), x1 + u[:, 2] * (x2 - x1), y2, x2,
                  y1 + u[:, 3] * (y2 - y1), (x1 + x2) / 2, (y1 + y2) / 2], 1)
    return torch.from_numpy(e.astype(np.float32))


def contours(rng, b, n=36):
    """36 clockwise contour points starting at the top (datasets/pipelines/loading.py:408-441 conventions).
    Returns raw (G, n, 2) points, the (G, 2n+2) table with the extent centre appended, and the extent boxes."""
    polys = []
    for g in range(len(b)):
        x1, y1, x2, y2 = b[g].numpy()
        cx, cy = (x1 + x2) / 2, (y1 + y2) / 2
        th = -np.pi / 2 + 2 * np.pi * np.arange(n) / n
        r = 0.6 + 0.4 * rng.rand(n)
        polys.append(np.stack([cx + r * (x2 - x1) / 2 * np.cos(th), cy + r * (y2 - y1) / 2 * np.sin(th)], 1))
    P = torch.from_numpy(np.stack(polys).astype(np.float32))
    xmin, ymin = P[:, :, 0].min(1)[0], P[:, :, 1].min(1)[0]
    xmax, ymax = P[:, :, 0].max(1)[0], P[:, :, 1].max(1)[0]
    ct = torch.stack([(xmin + xmax) / 2, (ymin + ymax) / 2], 1).unsqueeze(1)
    return P, torch.cat([P, ct], 1).reshape(len(b), -1), torch.stack([xmin, ymin, xmax, ymax], 1)


def keypoints(rng, b, n=17):
    G = len(b)
    x1, y1, x2, y2 = b.numpy().T
    x = x1[:, None] + rng.rand(G, n) * (x2 - x1)[:, None]
    y = y1[:, None] + rng.rand(G, n) * (y2 - y1)[:, None]
    v = rng.choice([0, 1, 2], size=(G, n), p=[0.2, 0.3, 0.5]).astype(np.float32)
    x, y = x * (v > 0), y * (v > 0)
    return torch.from_numpy(np.stack([x, y, v], 2).reshape(G, 3 * n).astype(np.float32))


def detector_batch(task, seed, B=2, H=384, W=384, max_gt=6, mixed_pad=True):
    rng = np.random.RandomState(seed)
    img = torch.from_numpy(rng.rand(B, 3, H, W).astype(np.float32))
    gt_b = [boxes(rng, rng.randint(1, max_gt + 1), H, W) for _ in range(B)]
    gt_l = [torch.from_numpy(rng.randint(0, 80, len(b))) for b in gt_b]
    metas = [dict(img_shape=(H, W, 3), pad_shape=(H, W, 3), scale_factor=1.0, flip=False) for _ in range(B)]
    if mixed_pad and B > 1:     # P13: per-image pad_shape smaller than the batch canvas
        metas[1] = dict(img_shape=(H - 40, W - 70, 3), pad_shape=(H - 32, W - 64, 3), scale_factor=1.0, flip=False)
    out = dict(img=img, gt_bboxes=gt_b, gt_labels=gt_l, img_metas=metas)
    if task == 'bbox':
        out['gt_extremes'] = [extremes(rng, b) for b in gt_b]
    elif task == 'segm':
        cs = [contours(rng, b) for b in gt_b]
        out['raw_polys'] = [c[0] for c in cs]
        out['gt_polygons'] = [c[1] for c in cs]
        out['gt_bboxes'] = [c[2] for c in cs]
        out['src_bboxes'] = gt_b
    elif task == 'pose_bbox':
        out['gt_keypoints_vs'] = [keypoints(rng, b) for b in gt_b]
        out['gt_labels'] = [torch.zeros(len(b), dtype=torch.long) for b in gt_b]
    return out


def loss_rows(loss_type, seed, N=64):
    """Random positive rows for cross_iou_loss: directional targets with exactly one selected slot per pair."""
    rng = np.random.RandomState(seed)
    NP = {'bbox': 5, 'polygon': 37, 'keypoint': 18}[loss_type]
    D = 4 * NP
    pred = torch.from_numpy((rng.rand(N, D) * 3 + 0.05).astype(np.float32))
    anchor = torch.from_numpy((rng.rand(N, 2) * 20).astype(np.float32))
    gt = torch.from_numpy(((rng.rand(N, 2 * NP) - 0.5) * 12).astype(np.float32)) + anchor.repeat(1, NP)
    weight = torch.from_numpy((rng.rand(N, 1) > 0.3).astype(np.float32)).repeat(1, D)
    bbox_gt = torch.from_numpy(np.concatenate([anchor.numpy() - rng.rand(N, 2) * 5 - 0.5,
                                               anchor.numpy() + rng.rand(N, 2) * 5 + 0.5], 1).astype(np.float32))
    vs = torch.from_numpy(rng.choice([0., 1., 2.], size=(N, NP - 1)).astype(np.float32))
    return dict(pred=pred, anchor=anchor, gt=gt, weight=weight, bbox_gt=bbox_gt, vs=vs, NP=NP, D=D)


def assign_case(seed, H=384, W=448, G=7, strides=(8, 16, 32, 64, 128)):
    """Points of a 5-level pyramid + GT boxes + jittered 'predicted' boxes for ATSS."""
    rng = np.random.RandomState(seed)
    sizes = [(int(np.ceil(H / s)), int(np.ceil(W / s))) for s in strides]
    gt = boxes(rng, G, H, W)
    pts = []
    for (h, w), s in zip(sizes, strides):
        xs = torch.arange(0., w) * s
        ys = torch.arange(0., h) * s
        pts.append(torch.stack([xs.repeat(h), ys.view(-1, 1).repeat(1, w).view(-1), xs.new_full((h * w,), s)], -1))
    flat = torch.cat(pts)
    half = torch.from_numpy((rng.rand(flat.shape[0], 4) * 3 + 0.2).astype(np.float32)) * flat[:, 2:3]
    jit = torch.from_numpy(((rng.rand(flat.shape[0], 2) - 0.5)).astype(np.float32)) * flat[:, 2:3]
    ctr = flat[:, :2] + jit
    pred = torch.cat([ctr - half[:, :2], ctr + half[:, 2:]], 1)
    return dict(sizes=sizes, strides=strides, points=flat, gt=gt, pred=pred,
                num_level=[h * w for h, w in sizes])
