"""ORACLE — test infrastructure only (see oracle/__init__.py).

Pure-PyTorch CPU restatement of the reference's LSNet training-step hot path, written functionally over a
``state_dict`` that uses the reference's parameter names.  Every function cites the reference
lines it follows (paths relative to /root/reference/code).  It is validated against the *unmodified* reference
(imported through ``ref_harness``) and pinned by the golden vectors under tests/golden/ — see
tests/golden/make_golden.py.  It is also the ``cpu_baseline`` / ``--impl reference`` arm of bench.py (kind "port"),
because the reference tree cannot travel to the GPU box.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import dcn_ops

INF = 1e8

# loss_weight of every term in the shipped configs (configs/lsnet/lsnet_{bbox,segm,pose_bbox}_r50_fpn_1x_coco.py)
DEFAULT_LOSS_WEIGHTS = {
    'bbox': {'loss_cls': 1.0, 'loss_bbox_init': 1.0, 'loss_bbox_refine': 2.0},
    'segm': {'loss_cls': 1.0, 'loss_segm_init': 1.0, 'loss_segm_refine': 2.0},
    'pose_bbox': {'loss_cls': 1.0, 'loss_bbox_init': 0.1, 'loss_bbox_refine': 0.2, 'loss_pose_init': 1.0,
                  'loss_pose_refine': 2.0},
}


# =====================================================================================================
# losses
# =====================================================================================================
def _bbox_from_extreme(pred, anchor_pts):
    """models/losses/cross_iou_loss.py:10-33."""
    pr = pred.view(pred.shape[0], -1, 2)
    val, ind = pr.max(dim=2)
    val = torch.where(ind == 0, -val, val)              # '-' slot wins ties and is negated
    val = val.view(val.shape[0], -1, 2)                 # (N, 5, [y, x])
    xs = val[:, :, 1] + anchor_pts[:, 0:1]
    ys = val[:, :, 0] + anchor_pts[:, 1:2]
    return torch.stack([xs[:, 1], ys[:, 0], xs[:, 3], ys[:, 2]], dim=1)


def _bbox_from_polygon(pred, anchor_pts):
    """models/losses/cross_iou_loss.py:35-59 (the trailing centre landmark is excluded)."""
    pr = pred[:, :-4].reshape(pred.shape[0], -1, 2)
    val, ind = pr.max(dim=2)
    val = torch.where(ind == 0, -val, val).view(pred.shape[0], -1, 2)
    xs = val[:, :, 1] + anchor_pts[:, 0:1]
    ys = val[:, :, 0] + anchor_pts[:, 1:2]
    return torch.stack([xs.min(1)[0], ys.min(1)[0], xs.max(1)[0], ys.max(1)[0]], dim=1)


def cross_iou_rows(pred, target, pos_inds, loss_type='bbox', anchor_pts=None, bbox_gt=None, vs=None, eps=1e-6,
                   alpha=0.2, stride=9):
    """Un-reduced per-row loss: models/losses/cross_iou_loss.py:61-132."""
    neg = ~pos_inds
    target = target.clone()
    target[neg] = alpha * target[pos_inds]                                   # :65-66
    if loss_type == 'polygon':
        tot = torch.stack([pred, target], -1).reshape(pred.size(0), -1, 4, 2)
        ov = []
        for i in range(stride):
            g = tot[:, i::stride].reshape(pred.size(0), -1, 2)
            ov.append(g.min(dim=2)[0].sum(1) / g.max(dim=2)[0].sum(1))
        overlaps = torch.stack(ov, -1).sum(-1) / stride
    elif loss_type == 'bbox':
        tot = torch.stack([pred, target], -1)
        overlaps = tot.min(dim=2)[0].sum(1) / tot.max(dim=2)[0].sum(1)
    else:
        tot = torch.stack([pred.reshape(pred.size(0), -1, 2), target.reshape(target.size(0), -1, 2)], -1)
        l_max = tot.max(dim=-1)[0].clamp(min=eps)
        l_min = tot.min(dim=-1)[0]
        overlaps = l_min.sum(-1) / l_max.sum(-1)
        vflag = (vs > 0).to(pred.dtype)
        vstack = torch.stack((vflag, vflag), 2).reshape(vs.size(0), -1)
        overlaps = torch.cat([overlaps[:, :-2] * vstack, overlaps[:, -2:]], 1)
        overlaps = overlaps.sum(-1) / tot.size(1)
    if loss_type == 'keypoint':
        return 1 - overlaps
    box = _bbox_from_extreme(pred, anchor_pts) if loss_type == 'bbox' else _bbox_from_polygon(pred, anchor_pts)
    enc_lt = torch.min(box[:, :2], bbox_gt[:, :2])
    enc_rb = torch.max(box[:, 2:], bbox_gt[:, 2:])
    enc = (enc_rb - enc_lt).clamp(min=0)
    c2 = enc[:, 0] ** 2 + enc[:, 1] ** 2 + eps
    w1, h1 = box[:, 2] - box[:, 0], box[:, 3] - box[:, 1] + eps
    w2, h2 = bbox_gt[:, 2] - bbox_gt[:, 0], bbox_gt[:, 3] - bbox_gt[:, 1] + eps
    rho2 = ((bbox_gt[:, 0] + bbox_gt[:, 2]) - (box[:, 0] + box[:, 2])) ** 2 / 4 + \
           ((bbox_gt[:, 1] + bbox_gt[:, 3]) - (box[:, 1] + box[:, 3])) ** 2 / 4
    v = (4 / math.pi ** 2) * torch.pow(torch.atan(w2 / h2) - torch.atan(w1 / h1), 2)
    return 1 - (overlaps - (rho2 / c2 + v ** 2 / (1 - overlaps + v)))


def cross_iou_loss(pred, target, weight, avg_factor, loss_weight=1.0, **kw):
    """CrossIOULoss.forward with reduction='mean' (cross_iou_loss.py:146-172; losses/utils.py:26-52)."""
    if weight is not None and not torch.any(weight > 0):
        return (pred * weight).sum()
    w = weight.mean(-1) if weight is not None and weight.dim() > 1 else weight
    loss = cross_iou_rows(pred, target, **kw)
    if w is not None:
        loss = loss * w
    return loss_weight * loss.sum() / avg_factor


def focal_loss(cls_score, labels, label_weights, avg_factor, gamma=2.0, alpha=0.25, loss_weight=1.0):
    """FocalLoss.forward -> sigmoid_focal_loss (models/losses/focal_loss.py:74-116,150-186)."""
    loss = dcn_ops.sigmoid_focal_loss_elementwise(cls_score, labels, gamma, alpha)
    loss = loss * label_weights.view(-1, 1)
    return loss_weight * loss.sum() / avg_factor


# =====================================================================================================
# assignment
# =====================================================================================================
def bbox_overlaps(b1, b2, eps=1e-6):
    """core/bbox/iou_calculators/iou2d_calculator.py:82-130 (mode='iou', not aligned)."""
    lt = torch.max(b1[:, None, :2], b2[:, :2])
    rb = torch.min(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[:, :, 0] * wh[:, :, 1]
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    union = torch.max(a1[:, None] + a2 - overlap, overlap.new_tensor([eps]))
    return overlap / union


def centroid_assign(points, gt_bboxes, scale=4, pos_num=1):
    """CentroidAssigner.assign, iou_type='center' (core/bbox/assigners/centroid_assigner.py:26-93).
    Returns 1-based gt index per point (0 = background)."""
    P, G = points.shape[0], gt_bboxes.shape[0]
    out = points.new_zeros((P,), dtype=torch.long)
    if P == 0 or G == 0:
        return out
    lvl = torch.log2(points[:, 2]).int()
    ctr = (gt_bboxes[:, :2] + gt_bboxes[:, 2:]) / 2
    wh = (gt_bboxes[:, 2:] - gt_bboxes[:, :2]).clamp(min=1e-6)
    glvl = ((torch.log2(wh[:, 0] / scale) + torch.log2(wh[:, 1] / scale)) / 2).int()
    glvl = torch.clamp(glvl, min=lvl.min(), max=lvl.max())
    dist = ((points[:, None, :2] - ctr[None]) / wh[None]).norm(dim=2)
    dist[lvl[:, None] != glvl[None, :]] = INF
    md, mi = torch.topk(dist, pos_num, dim=0, largest=False)
    dinf = torch.full_like(dist, INF)
    dinf[mi, torch.arange(G)] = md
    md, mi = dinf.min(dim=1)
    out[md != INF] = mi[md != INF] + 1
    return out


def atss_assign(bboxes, num_level, gt_bboxes, topk=9):
    """ATSSAssigner.assign (core/bbox/assigners/atss_assigner.py:29-164).  Returns (gt_inds 1-based, max_overlaps)."""
    NEG = -100000000
    P, G = bboxes.shape[0], gt_bboxes.shape[0]
    ov = bbox_overlaps(bboxes, gt_bboxes)
    out = ov.new_full((P,), 0, dtype=torch.long)
    if P == 0 or G == 0:
        return out, ov.new_zeros((P,))
    gcx, gcy = (gt_bboxes[:, 0] + gt_bboxes[:, 2]) / 2.0, (gt_bboxes[:, 1] + gt_bboxes[:, 3]) / 2.0
    bcx, bcy = (bboxes[:, 0] + bboxes[:, 2]) / 2.0, (bboxes[:, 1] + bboxes[:, 3]) / 2.0
    dist = (torch.stack((bcx, bcy), 1)[:, None] - torch.stack((gcx, gcy), 1)[None]).pow(2).sum(-1).sqrt()
    cands, s = [], 0
    for n in num_level:
        _, idx = dist[s:s + n].topk(topk, dim=0, largest=False)
        cands.append(idx + s)
        s += n
    cands = torch.cat(cands, 0)                                          # (topk*L, G)
    cov = ov[cands, torch.arange(G)]
    thr = cov.mean(0) + cov.std(0)
    is_pos = cov >= thr[None]
    l_ = bcx[cands] - gt_bboxes[:, 0]
    t_ = bcy[cands] - gt_bboxes[:, 1]
    r_ = gt_bboxes[:, 2] - bcx[cands]
    b_ = gt_bboxes[:, 3] - bcy[cands]
    is_pos = is_pos & (torch.stack([l_, t_, r_, b_], 1).min(1)[0] > 0.01)
    ov_inf = torch.full_like(ov, NEG)
    gsel = torch.arange(G)[None].expand_as(cands)
    ov_inf[cands[is_pos], gsel[is_pos]] = ov[cands[is_pos], gsel[is_pos]]
    mo, am = ov_inf.max(dim=1)
    out[mo != NEG] = am[mo != NEG] + 1
    return out, mo


# =====================================================================================================
# target generation + loss (LSHead.loss)
# =====================================================================================================
def grid_points(h, w, stride):
    """PointGenerator.grid_points (core/anchor/point_generator.py:17-25): rows [x, y, stride], row-major."""
    xs = torch.arange(0., w) * stride
    ys = torch.arange(0., h) * stride
    return torch.stack([xs.repeat(h), ys.view(-1, 1).repeat(1, w).view(-1), xs.new_full((h * w,), stride)], -1)


def valid_flags(h, w, vh, vw):
    """PointGenerator.valid_flags (point_generator.py:27-37)."""
    fx = torch.zeros(w, dtype=torch.bool); fx[:vw] = True
    fy = torch.zeros(h, dtype=torch.bool); fy[:vh] = True
    return fx.repeat(h) & fy.view(-1, 1).repeat(1, w).view(-1)


def directional_targets(gt_pts, anchor_pts, weights):
    """LSHead.get_bbox_gt_reg / get_poly_gt_reg (dense_heads/lsnet_head.py:402-454).
    gt_pts (N, 2*NP) as (x,y) pairs -> targets (N, 4*NP) [y-,y+,x-,x+] per landmark + bool slot mask."""
    N, NP = gt_pts.shape[0], gt_pts.shape[1] // 2
    off = (gt_pts - anchor_pts[:, :2].repeat(1, NP)).view(N, NP, 2)       # (x, y)
    pos = off >= 0
    mag = off.abs() * (weights[:, :1] > 0).to(off.dtype).view(N, 1, 1)
    zero = torch.zeros_like(mag)
    # per landmark [y-, y+, x-, x+]
    tgt = torch.stack([torch.where(pos[..., 1], zero[..., 1], mag[..., 1]),
                       torch.where(pos[..., 1], mag[..., 1], zero[..., 1]),
                       torch.where(pos[..., 0], zero[..., 0], mag[..., 0]),
                       torch.where(pos[..., 0], mag[..., 0], zero[..., 0])], -1).reshape(N, 4 * NP)
    sel = torch.stack([~pos[..., 1], pos[..., 1], ~pos[..., 0], pos[..., 0]], -1).reshape(N, 4 * NP)
    return tgt, sel


def extreme_points2bbox(pts):
    """LSHead.extreme_points2bbox (lsnet_head.py:321-347), pts (B, 20, H, W) -> (B, 4, H, W)."""
    pr = pts.view(pts.shape[0], -1, 2, *pts.shape[2:])
    val, ind = pr.max(dim=2)
    val = torch.where(ind == 0, -val, val)
    val = val.view(val.shape[0], -1, 2, *val.shape[2:])
    ys, xs = val[:, :, 0], val[:, :, 1]
    return torch.stack([xs[:, 1], ys[:, 0], xs[:, 3], ys[:, 2]], 1)


def vectors2bbox(pts):
    """LSHead.vectors2bbox (lsnet_head.py:349-370)."""
    pr = pts[:, :-4].reshape(pts.shape[0], -1, 2, *pts.shape[2:])
    val, ind = pr.max(dim=2)
    val = torch.where(ind == 0, -val, val)
    val = val.view(val.shape[0], -1, 2, *val.shape[2:])
    ys, xs = val[:, :, 0], val[:, :, 1]
    return torch.stack([xs.min(1)[0], ys.min(1)[0], xs.max(1)[0], ys.max(1)[0]], 1)


def border_center(gt_bboxes):
    """LSHead.get_border_center (lsnet_head.py:1677-1697): (G,10) [top, left, bottom, right, centre] as (x,y)."""
    x1, y1, x2, y2 = gt_bboxes.unbind(1)
    cx, cy = (x2 + x1) / 2.0, (y2 + y1) / 2.0
    return torch.stack([cx, y1, x1, cy, cx, y2, x2, cy, cx, cy], 1)


def process_keypoints_with_bbox(gt_bboxes, gt_kps_vs):
    """LSHead.process_keypoints_with_bbox (lsnet_head.py:1758-1785)."""
    x, y, v = gt_kps_vs[:, 0::3], gt_kps_vs[:, 1::3], gt_kps_vs[:, 2::3]
    ct = torch.stack([(gt_bboxes[:, 0] + gt_bboxes[:, 2]) / 2, (gt_bboxes[:, 1] + gt_bboxes[:, 3]) / 2], 1)
    return torch.cat((torch.stack((x, y), 2).reshape(gt_kps_vs.size(0), -1), ct), 1), v


def target_single(stage, flat, valid, num_level, gt_bboxes, gt_pts, gt_vs, gt_labels, num_classes, cfg):
    """LSHead._target_single (lsnet_head.py:796-917) for one image.  gt_pts: (G, 2*NP) landmark table of the task."""
    props = flat[valid]
    if stage == 'init':
        a = centroid_assign(props, gt_bboxes, cfg.get('scale', 4), cfg.get('pos_num', 1))
    else:
        inside = [int(f.sum()) for f in torch.split(valid, num_level)]
        a, _ = atss_assign(props[:, :4], inside, gt_bboxes, cfg.get('topk', 9))
    pos = torch.nonzero(a > 0, as_tuple=False).squeeze(-1).unique()
    neg = torch.nonzero(a == 0, as_tuple=False).squeeze(-1).unique()
    n = props.shape[0]
    bboxes_gt = props.new_zeros((n, 4))
    pts_gt = props.new_zeros((n, gt_pts.shape[1]))
    vs_gt = props.new_zeros((n, gt_vs.shape[1])) if gt_vs is not None else None
    weights = props.new_zeros((n, 4))
    labels = props.new_full((n,), num_classes, dtype=torch.long)
    lw = props.new_zeros(n)
    if len(pos) > 0:
        gi = a[pos] - 1
        bboxes_gt[pos] = gt_bboxes[gi]
        pts_gt[pos] = gt_pts[gi]
        if vs_gt is not None:
            vs_gt[pos] = gt_vs[gi]
        weights[pos] = 1.0
        labels[pos] = gt_labels[gi] if gt_labels is not None else 1
        lw[pos] = 1.0
    if len(neg) > 0:
        lw[neg] = 1.0

    def unmap(d, fill=0):
        out = d.new_full((flat.shape[0],) + d.shape[1:], fill)
        out[valid] = d
        return out
    return dict(labels=unmap(labels, num_classes), label_weights=unmap(lw), bboxes_gt=unmap(bboxes_gt),
                pts_gt=unmap(pts_gt), vs_gt=unmap(vs_gt) if vs_gt is not None else None, weights=unmap(weights),
                npos=max(pos.numel(), 1), assign=unmap(a, 0))


def head_loss(outs, gt_bboxes, gt_labels, img_metas, task='bbox', gt_extremes=None, gt_polygons=None,
              gt_keypoints_vs=None, num_classes=80, strides=(8, 16, 32, 64, 128), base_scale=4,
              loss_weights=None, return_aux=False):
    """LSHead.loss (lsnet_head.py:1272-1437) for task in {'bbox', 'segm', 'pose_bbox'}.
    outs: dict with per-level lists 'cls', 'bbox_init', 'bbox_refine' (+ 'segm_*' / 'pose_*')."""
    cls_scores = outs['cls']
    lwt = dict(DEFAULT_LOSS_WEIGHTS[task])
    lwt.update(loss_weights or {})
    B, L = cls_scores[0].shape[0], len(cls_scores)
    sizes = [c.shape[-2:] for c in cls_scores]
    num_level = [int(h * w) for h, w in sizes]
    gt_vs = [None] * B
    if task == 'bbox':
        if gt_extremes is None:
            gt_extremes = [border_center(b) for b in gt_bboxes]
        gt_pts, prim = gt_extremes, 'bbox'
    elif task == 'segm':
        gt_pts, prim = gt_polygons, 'segm'           # (G, 74) incl. centre; gt_bboxes = polygon extents
    else:
        if gt_extremes is None:
            gt_extremes = [border_center(b) for b in gt_bboxes]
        kp = [process_keypoints_with_bbox(b, k) for b, k in zip(gt_bboxes, gt_keypoints_vs)]
        gt_kps, gt_vs = [k[0] for k in kp], [k[1] for k in kp]
        gt_pts, prim = gt_extremes, 'bbox'
    pts = [grid_points(h, w, s) for (h, w), s in zip(sizes, strides)]
    flat_pts = torch.cat(pts)
    valids = []
    for m in img_metas:
        ph, pw = m['pad_shape'][:2]
        valids.append(torch.cat([valid_flags(h, w, min(int(np.ceil(ph / s)), h), min(int(np.ceil(pw / s)), w))
                                 for (h, w), s in zip(sizes, strides)]))
    init_pred = outs[prim + '_init']
    to_box = vectors2bbox if task == 'segm' else extreme_points2bbox
    tg = {'init': [], 'refine': []}
    for i in range(B):
        extra = {}
        if task == 'pose_bbox':
            # both landmark tables ride along: extremes (10) ++ keypoints (36)
            table = torch.cat([gt_pts[i], gt_kps[i]], 1)
        else:
            table = gt_pts[i]
        tg['init'].append(target_single('init', flat_pts, valids[i], num_level, gt_bboxes[i], table, gt_vs[i],
                                        gt_labels[i], num_classes, dict(scale=4, pos_num=1)))
        boxes = []
        for l in range(L):
            shift = to_box(init_pred[l].detach())[i] * strides[l]
            ctr = torch.cat([pts[l][:, :2], pts[l][:, :2]], 1)
            boxes.append(ctr + shift.permute(1, 2, 0).reshape(-1, 4))
        tg['refine'].append(target_single('refine', torch.cat(boxes), valids[i], num_level, gt_bboxes[i], table,
                                          gt_vs[i], gt_labels[i], num_classes, dict(topk=9)))
    npos = {k: sum(t['npos'] for t in v) for k, v in tg.items()}

    def per_level(stage, key):
        full = torch.stack([t[key] for t in tg[stage]], 0)
        out, s = [], 0
        for n in num_level:
            out.append(full[:, s:s + n])
            s += n
        return out
    labels, lweights = per_level('refine', 'labels'), per_level('refine', 'label_weights')
    names = {'bbox': ['loss_bbox_init', 'loss_bbox_refine'], 'segm': ['loss_segm_init', 'loss_segm_refine'],
             'pose_bbox': ['loss_bbox_init', 'loss_bbox_refine', 'loss_pose_init', 'loss_pose_refine']}[task]
    losses = {'loss_cls': []}
    for nme in names:
        losses[nme] = []
    for l in range(L):
        s = strides[l]
        cs = cls_scores[l].permute(0, 2, 3, 1).reshape(-1, num_classes)
        losses['loss_cls'].append(focal_loss(cs, labels[l].reshape(-1), lweights[l].reshape(-1), npos['refine'],
                                             loss_weight=lwt['loss_cls']))
        anchor = torch.stack([pts[l]] * B, 0).reshape(-1, 3)
        nt = base_scale * s
        for stage in ('init', 'refine'):
            bg = per_level(stage, 'bboxes_gt')[l].reshape(-1, 4)
            pg = per_level(stage, 'pts_gt')[l].reshape(B * num_level[l], -1)
            w4 = per_level(stage, 'weights')[l].reshape(-1, 4)
            if task in ('bbox', 'pose_bbox'):
                ext = pg[:, :10]
                wt = w4.repeat(1, 5)
                pred = outs['bbox_' + stage][l].permute(0, 2, 3, 1).reshape(-1, 20) * s
                t, sel = directional_targets(ext, anchor, wt)
                losses['loss_bbox_' + stage].append(cross_iou_loss(
                    pred / nt, t / nt, wt, npos[stage], lwt['loss_bbox_' + stage], loss_type='bbox', anchor_pts=anchor[:, :-1] / nt,
                    bbox_gt=bg / nt, pos_inds=sel))
            if task == 'segm':
                D = pg.shape[1] * 2
                wt = w4[:, :1].repeat(1, D)
                pred = outs['segm_' + stage][l].permute(0, 2, 3, 1).reshape(-1, D) * s
                t, sel = directional_targets(pg, anchor, wt)
                losses['loss_segm_' + stage].append(cross_iou_loss(
                    pred / nt, t / nt, wt, npos[stage], lwt['loss_segm_' + stage], loss_type='polygon', anchor_pts=anchor[:, :-1] / nt,
                    bbox_gt=bg / nt, pos_inds=sel))
            if task == 'pose_bbox':
                kpg = pg[:, 10:]
                D = kpg.shape[1] * 2
                wt = w4[:, :1].repeat(1, D)
                vsg = per_level(stage, 'vs_gt')[l].reshape(B * num_level[l], -1)
                pred = outs['pose_' + stage][l].permute(0, 2, 3, 1).reshape(-1, D) * s
                t, sel = directional_targets(kpg, anchor, wt)
                losses['loss_pose_' + stage].append(cross_iou_loss(
                    pred / nt, t / nt, wt, npos[stage], lwt['loss_pose_' + stage], loss_type='keypoint', anchor_pts=anchor[:, :-1] / nt,
                    bbox_gt=None, pos_inds=sel, vs=vsg.clone()))
    if return_aux:
        return losses, dict(tg=tg, npos=npos)
    return losses


# =====================================================================================================
# network forward (functional over the reference's state_dict keys)
# =====================================================================================================
def _gn(x, sd, prefix, groups=32):
    return F.group_norm(x, groups, sd[prefix + '.weight'], sd[prefix + '.bias'], 1e-5)


def _bn_eval(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'], sd[prefix + '.weight'],
                        sd[prefix + '.bias'], False, 0.0, 1e-5)


def resnet50_forward(sd, x, prefix='backbone.'):
    """ResNet-50, style='pytorch', BN in eval (models/backbones/resnet.py:261-301,619-646)."""
    x = F.relu(_bn_eval(F.conv2d(x, sd[prefix + 'conv1.weight'], None, 2, 3), sd, prefix + 'bn1'))
    x = F.max_pool2d(x, 3, 2, 1)
    outs = []
    for li, nblocks in enumerate((3, 4, 6, 3), 1):
        for bi in range(nblocks):
            p = f'{prefix}layer{li}.{bi}.'
            stride = 2 if (bi == 0 and li > 1) else 1
            idt = x
            o = F.relu(_bn_eval(F.conv2d(x, sd[p + 'conv1.weight']), sd, p + 'bn1'))
            o = F.relu(_bn_eval(F.conv2d(o, sd[p + 'conv2.weight'], None, stride, 1), sd, p + 'bn2'))
            o = _bn_eval(F.conv2d(o, sd[p + 'conv3.weight']), sd, p + 'bn3')
            if p + 'downsample.0.weight' in sd:
                idt = _bn_eval(F.conv2d(x, sd[p + 'downsample.0.weight'], None, stride), sd, p + 'downsample.1')
            x = F.relu(o + idt)
        outs.append(x)
    return outs


def fpn_forward(sd, feats, prefix='neck.'):
    """FPN(start_level=1, add_extra_convs='on_input', num_outs=5, GN) (models/necks/fpn.py:165-217)."""
    ins = feats[1:]
    lat = [_gn(F.conv2d(ins[i], sd[f'{prefix}lateral_convs.{i}.conv.weight']), sd, f'{prefix}lateral_convs.{i}.gn')
           for i in range(3)]
    for i in (2, 1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode='nearest')
    outs = [_gn(F.conv2d(lat[i], sd[f'{prefix}fpn_convs.{i}.conv.weight'], None, 1, 1), sd, f'{prefix}fpn_convs.{i}.gn')
            for i in range(3)]
    outs.append(_gn(F.conv2d(ins[-1], sd[f'{prefix}fpn_convs.3.conv.weight'], None, 2, 1), sd, f'{prefix}fpn_convs.3.gn'))
    outs.append(_gn(F.conv2d(outs[-1], sd[f'{prefix}fpn_convs.4.conv.weight'], None, 2, 1), sd, f'{prefix}fpn_convs.4.gn'))
    return outs


def _dcn_tower(sd, x, prefix, n=3):
    """3 x DCNConvModule = ModulatedDeformConvPack -> GN -> ReLU (lsnet_head.py:1830-1849; deform_conv.py:527-534)."""
    for i in range(n):
        p = f'{prefix}.{i}.'
        o = F.conv2d(x, sd[p + 'conv.conv_offset.weight'], sd[p + 'conv.conv_offset.bias'], 1, 1)
        o1, o2, m = torch.chunk(o, 3, dim=1)
        x = dcn_ops.modulated_deform_conv(x, torch.cat((o1, o2), 1), torch.sigmoid(m), sd[p + 'conv.weight'],
                                          sd[p + 'conv.bias'], 1, 1, 1, 1, 1)
        x = F.relu(_gn(x, sd, p + 'bn'))
    return x


def _pred_reg(sp, free=None, task='bbox', num_vectors=4):
    """LSHead.get_pred_reg (lsnet_head.py:372-400): 9 signed (y,x) DCN sampling points."""
    if free is not None:
        pr = sp.view(sp.shape[0], -1, 2, *sp.shape[2:])
        val, ind = pr.max(dim=2)
        return torch.cat((torch.where(ind == 0, -val, val), free), 1)
    r = sp.view(sp.shape[0], -1, 4, *sp.shape[2:])
    ct, poly = r[:, -1:], r[:, :-1]
    if task == 'segm':
        sel = poly[:, ::math.ceil(num_vectors / 8)]
    else:
        sel = poly[:, 1::2]
    offs = torch.cat([sel, ct], 1)
    offs = offs.reshape(offs.shape[0], -1, 2, *offs.shape[3:])
    val, ind = offs.max(dim=2)
    return torch.where(ind == 0, -val, val)


def _pyramid_dcn(x, offset, weight, sh, sw):
    """PyramidDeformConv.forward incl. the tiny-input zero pad (ops/dcn/deform_conv.py:611-630)."""
    ph, pw = max(3 - x.size(2), 0), max(3 - x.size(3), 0)
    if ph or pw:
        x = F.pad(x, (0, pw, 0, ph))
        offset = F.pad(offset, (0, pw, 0, ph))
    out = dcn_ops.pyramid_deform_conv(x, offset, weight, (sh, sw), 1, 1, 1, 1, 1)
    if ph or pw:
        out = out[:, :, :out.size(2) - ph, :out.size(3) - pw].contiguous()
    return out


def head_forward(sd, feats, task='bbox', num_vectors=4, gradient_mul=0.1, prefix='bbox_head.'):
    """LSHead.forward (lsnet_head.py:479-755), conv_module_type='dcn'.  Returns dict of per-level lists."""
    p = prefix
    base = torch.tensor([[y, x] for y in (-1., 0., 1.) for x in (-1., 0., 1.)]).view(1, 18, 1, 1).to(feats[0].dtype)
    branches = {'bbox': ['bbox'], 'segm': ['segm'], 'pose_bbox': ['bbox', 'pose']}[task]
    cls_feats = [_dcn_tower(sd, f, p + 'cls_convs') for f in feats]
    bf, init_sp, dcn_off = {}, {}, {}
    for br in branches:
        bf[br] = [_dcn_tower(sd, f, p + br + '_convs') for f in feats]
        init_sp[br], dcn_off[br] = [], []
        for f in bf[br]:
            o = F.conv2d(F.relu(F.conv2d(f, sd[f'{p}pts_{br}_init_conv.weight'], sd[f'{p}pts_{br}_init_conv.bias'], 1, 1)),
                         sd[f'{p}pts_{br}_init_out.weight'], sd[f'{p}pts_{br}_init_out.bias'])
            if br == 'bbox':
                sp = F.softplus(o[:, :20])
                reg = _pred_reg(sp, o[:, 20:])
            else:
                sp = F.softplus(o)
                reg = _pred_reg(sp, None, task, num_vectors)
            reg = (1 - gradient_mul) * reg.detach() + gradient_mul * reg
            init_sp[br].append(sp)
            dcn_off[br].append(reg - base)
    L = len(feats)
    outs = {'cls': []}
    for br in branches:
        outs[br + '_init'] = init_sp[br]
        outs[br + '_refine'] = []
    cls_driver = branches[-1]          # pts_cls_conv follows the pose offsets when both exist (lsnet_head.py:680-681)
    for l in range(L):
        lvls = [l, l + 1, l + 2] if l == 0 else ([l, l - 1, l - 2] if l == L - 1 else [l, l - 1, l + 1])
        bh, bw = cls_feats[l].shape[2:]
        raws = {br: [] for br in branches}
        cls_raws = []
        offs = {br: dcn_off[br][l] for br in branches}
        for lv in lvls:
            sh, sw = cls_feats[lv].size(2) / bh, cls_feats[lv].size(3) / bw
            for br in branches:
                # in-place scaling on views => CUMULATIVE over the three iterations (lsnet_head.py:628-633, trap P1)
                oy = offs[br][:, 0::2] * sh
                ox = offs[br][:, 1::2] * sw
                offs[br] = torch.stack([oy, ox], 2).view(oy.size(0), -1, oy.size(2), oy.size(3))
                raws[br].append(_pyramid_dcn(bf[br][lv], offs[br], sd[f'{p}pts_{br}_refine_conv.weight'], sh, sw))
            cls_raws.append(_pyramid_dcn(cls_feats[lv], offs[cls_driver], sd[p + 'pts_cls_conv.weight'], sh, sw))
        for br in branches:
            t = F.relu(F.conv2d(torch.cat(raws[br], 1), sd[f'{p}{br}_af_dcn_conv.0.weight'],
                                sd[f'{p}{br}_af_dcn_conv.0.bias']))
            t = t + F.conv2d(bf[br][l], sd[f'{p}{br}_feat_conv.weight'], sd[f'{p}{br}_feat_conv.bias'], 1, 1)
            t = F.conv2d(F.relu(_gn(t, sd, f'{p}{br}_GN')), sd[f'{p}pts_{br}_refine_out.weight'],
                         sd[f'{p}pts_{br}_refine_out.bias'])
            outs[br + '_refine'].append(F.softplus(t + init_sp[br][l].detach()))
        t = F.relu(F.conv2d(torch.cat(cls_raws, 1), sd[p + 'cls_af_dcn_conv.0.weight'], sd[p + 'cls_af_dcn_conv.0.bias']))
        t = t + F.conv2d(cls_feats[l], sd[p + 'cls_feat_conv.weight'], sd[p + 'cls_feat_conv.bias'], 1, 1)
        outs['cls'].append(F.conv2d(F.relu(_gn(t, sd, p + 'cls_GN')), sd[p + 'pts_cls_out.weight'],
                                    sd[p + 'pts_cls_out.bias']))
    return outs


def detector_losses(sd, img, gt_bboxes, gt_labels, img_metas, task='bbox', **kw):
    """LSDetector.forward_train (models/detectors/lsnet.py:44-56): backbone -> FPN -> head -> loss dict."""
    feats = fpn_forward(sd, resnet50_forward(sd, img))
    outs = head_forward(sd, feats, task=task, num_vectors={'bbox': 4, 'segm': 36, 'pose_bbox': 17}[task])
    nc = kw.pop('num_classes', 1 if task == 'pose_bbox' else 80)
    return head_loss(outs, gt_bboxes, gt_labels, img_metas, task=task, num_classes=nc, **kw)


def parse_losses(losses):
    """BaseDetector._parse_losses (models/detectors/base.py:176-209), single process."""
    log = {k: sum(v) if isinstance(v, (list, tuple)) else v for k, v in losses.items()}
    return sum(v for k, v in log.items() if 'loss' in k), log


def trainable_keys(sd):
    """Frozen: stem + stage 1 (frozen_stages=1) and every BN (norm_eval) keeps its statistics; BN affine params of
    stages 2-4 remain trainable (requires_grad=True in the config) (resnet.py:569-585, 636-646)."""
    keys = []
    for k in sd:
        if 'running_' in k or 'num_batches_tracked' in k:
            continue
        if k.startswith('backbone.conv1') or k.startswith('backbone.bn1') or k.startswith('backbone.layer1.'):
            continue
        keys.append(k)
    return keys


def sgd_step(sd, grads, momentum_buf, lr=0.01, momentum=0.9, weight_decay=1e-4, max_norm=35.0):
    """clip_grad_norm_(35, L2) then torch.optim.SGD step (mmcv/runner/hooks/optimizer.py:19-28;
    configs/lsnet/lsnet_bbox_r50_fpn_1x_coco.py:64-65; configs/_base_/schedules/schedule_1x.py:2)."""
    total = torch.sqrt(sum((g.detach() ** 2).sum() for g in grads.values()))
    coef = (max_norm / (total + 1e-6)).clamp(max=1.0)
    with torch.no_grad():
        for k, g in grads.items():
            d = g * coef + weight_decay * sd[k]
            buf = momentum_buf.get(k)
            buf = d.clone() if buf is None else buf.mul_(momentum).add_(d)
            momentum_buf[k] = buf
            sd[k].add_(buf, alpha=-lr)
    return total
