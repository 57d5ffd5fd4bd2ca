"""ORACLE — test infrastructure only (see oracle/__init__.py).

CPU restatement of the reference's inference decode: ``LSHead.get_bboxes`` / ``_get_bboxes_single``
(mmdet/models/dense_heads/lsnet_head.py:1439-1668), ``extreme_points2bbox`` / ``vectors2bbox`` (:321-370),
``multiclass_nms_lsvr`` (mmdet/core/post_processing/bbox_nms.py:60-99), ``batched_nms`` / ``nms``
(mmdet/ops/nms/nms_wrapper.py:7-157; greedy suppression of mmdet/ops/nms/src/cpu/nms_cpu.cpp: sort by descending
score, drop every later box whose IoU with a kept box exceeds the threshold; areas without the +1).

Pinned against the reference's own Python run in this container (tests/golden/make_golden.py -> decode_*.npz; the
reference's compiled ``nms_ext`` is absent, ``greedy_nms`` below stands in for it there too, so the NMS core itself is
pinned by its published definition and the hand-checked example of nms_wrapper.py:25-34)."""
import numpy as np
import torch


def greedy_nms(dets, iou_thr):
    """dets (N,5) -> indices kept, in descending score order."""
    if dets.shape[0] == 0:
        return torch.zeros(0, dtype=torch.long)
    d = dets.detach().cpu().numpy().astype(np.float32)
    order = np.argsort(-d[:, 4], kind='stable')
    x1, y1, x2, y2 = d[:, 0], d[:, 1], d[:, 2], d[:, 3]
    area = (x2 - x1) * (y2 - y1)
    keep = []
    supp = np.zeros(len(d), bool)
    for ii, i in enumerate(order):
        if supp[i]:
            continue
        keep.append(i)
        rest = order[ii + 1:]
        w = np.maximum(np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]), np.float32(0))
        h = np.maximum(np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]), np.float32(0))
        inter = (w * h).astype(np.float32)
        iou = inter / (area[i] + area[rest] - inter)
        supp[rest[iou > np.float32(iou_thr)]] = True
    return torch.as_tensor(np.asarray(keep, dtype=np.int64))


def batched_nms(bboxes, scores, inds, iou_thr):
    max_coordinate = bboxes.max()
    offsets = inds.to(bboxes) * (max_coordinate + 1)
    boxes_for_nms = bboxes + offsets[:, None]
    keep = greedy_nms(torch.cat([boxes_for_nms, scores[:, None]], -1), iou_thr)
    return torch.cat([bboxes[keep], scores[keep][:, None]], -1), keep


def multiclass_nms_lsvr(multi_bboxes, multi_pts, multi_scores, npts, score_thr, iou_thr, max_num=-1):
    num_classes = multi_scores.size(1) - 1
    bboxes = multi_bboxes[:, None].expand(-1, num_classes, 4)
    pts = multi_pts[:, None].expand(-1, num_classes, multi_pts.shape[-1])
    scores = multi_scores[:, :-1]
    valid = scores > score_thr
    bboxes, pts, scores = bboxes[valid], pts[valid], scores[valid]
    labels = valid.nonzero()[:, 1]
    if bboxes.numel() == 0:
        return multi_bboxes.new_zeros((0, 5)), pts.new_zeros((0, npts * 2)), multi_bboxes.new_zeros((0,), dtype=torch.long)
    dets, keep = batched_nms(bboxes, scores, labels, iou_thr)
    if max_num > 0:
        dets, keep = dets[:max_num], keep[:max_num]
    return dets, pts[keep], labels[keep]


def _signed(pts):
    r = pts.view(pts.shape[0], -1, 2, *pts.shape[2:])
    val, ind = torch.max(r, dim=2)
    val = torch.where(ind == 0, -val, val)
    return val.view(val.shape[0], -1, 2, *val.shape[2:])


def extreme_points2bbox(pts):
    v = _signed(pts)
    py, px = v[:, :, 0], v[:, :, 1]
    bbox = torch.stack([px[:, 1], py[:, 0], px[:, 3], py[:, 2]], 1)
    ext = torch.stack([px[:, 0], py[:, 0], px[:, 1], py[:, 1], px[:, 2], py[:, 2], px[:, 3], py[:, 3]], 1)
    return ext, bbox


def vectors2bbox(pts):
    v = _signed(pts[:, :-4])
    py, px = v[:, :, 0], v[:, :, 1]
    bbox = torch.stack([px.min(1)[0], py.min(1)[0], px.max(1)[0], py.max(1)[0]], 1)
    vec = torch.stack([px, py], 2).reshape(py.shape[0], -1, *py.shape[2:])
    return vec, bbox


def get_bboxes(task, num_vectors, strides, cls_scores, bbox_refine, lm_refine, img_metas, test_cfg, rescale=False):
    """cls_scores[l] (B,C,H,W); bbox_refine[l] (B,20,H,W) for tasks with a box branch; lm_refine[l] (B,4(n+1),H,W) landmark
    branch (segm / pose).  Returns per image (det_bboxes, det_pts, det_labels)."""
    if task in ('bbox', 'pose_bbox'):
        ext = [extreme_points2bbox(p) for p in bbox_refine]
    if task in ('segm', 'pose_bbox', 'pose_kbox'):
        vec = [vectors2bbox(p) for p in lm_refine]
    box_src = ext if task in ('bbox', 'pose_bbox') else vec
    pts_src = ext if task == 'bbox' else vec
    nv = num_vectors
    out = []
    for b, meta in enumerate(img_metas):
        mb, mp, ms = [], [], []
        for l, s in enumerate(strides):
            C, h, w = cls_scores[l].shape[1:]
            xs = torch.arange(0., w) * s
            ys = torch.arange(0., h) * s
            points = torch.stack([xs.repeat(h), ys.view(-1, 1).repeat(1, w).view(-1)], -1)
            scores = cls_scores[l][b].permute(1, 2, 0).reshape(-1, C).sigmoid()
            bp = box_src[l][1][b].permute(1, 2, 0).reshape(-1, 4)
            pp = pts_src[l][0][b].permute(1, 2, 0).reshape(-1, nv * 2)
            nms_pre = test_cfg.get('nms_pre', -1)
            if nms_pre > 0 and scores.shape[0] > nms_pre:
                _, topk = scores.max(dim=1)[0].topk(nms_pre)
                points, bp, pp, scores = points[topk], bp[topk], pp[topk], scores[topk]
            bboxes = bp * s + torch.cat([points, points], 1)
            pts = pp * s + points.repeat(1, nv)
            H, W = meta['img_shape'][:2]
            x1, y1 = bboxes[:, 0].clamp(min=0, max=W), bboxes[:, 1].clamp(min=0, max=H)
            x2, y2 = bboxes[:, 2].clamp(min=0, max=W), bboxes[:, 3].clamp(min=0, max=H)
            mb.append(torch.stack([x1, y1, x2, y2], -1))
            if task == 'bbox':
                xt, yl = pts[:, 0].clamp(min=0, max=W), pts[:, 3].clamp(min=0, max=H)
                xb, yr = pts[:, 4].clamp(min=0, max=W), pts[:, 7].clamp(min=0, max=H)
                mp.append(torch.stack([xt, y1, x1, yl, xb, y2, x2, yr], -1))
            else:
                mp.append(torch.stack([pts[:, 0::2].clamp(min=0, max=W), pts[:, 1::2].clamp(min=0, max=H)], 2).reshape(pts.size(0), -1))
            ms.append(scores)
        mb, mp, ms = torch.cat(mb), torch.cat(mp), torch.cat(ms)
        if rescale:
            sf = np.asarray(meta['scale_factor'], np.float32)
            mb = mb / mb.new_tensor(sf)
            mp = mp / mp.new_tensor(np.tile(sf, 2) if task == 'bbox' else np.tile(sf[:2], nv))
        ms = torch.cat([ms, ms.new_zeros(ms.shape[0], 1)], 1)
        out.append(multiclass_nms_lsvr(mb, mp, ms, nv, test_cfg['score_thr'], test_cfg['nms']['iou_thr'], test_cfg['max_per_img']))
    return out


def synth_head_outputs(task, seed, B=2, sizes=((40, 52), (20, 26), (10, 13), (5, 7), (3, 4)), num_classes=None):
    """Seeded head outputs for decode tests: class logits around the score threshold, positive landmark slot pairs."""
    g = torch.Generator().manual_seed(seed)
    nv = {'bbox': 4, 'segm': 36, 'pose_bbox': 17}[task]
    C = num_classes or (1 if task == 'pose_bbox' else 80)
    cls = [torch.randn(B, C, h, w, generator=g) * 1.5 - (6.5 if C > 1 else 3.0) for h, w in sizes]
    box = [torch.rand(B, 20, h, w, generator=g) * 3 for h, w in sizes] if task in ('bbox', 'pose_bbox') else None
    lm = [torch.rand(B, 4 * (nv + 1), h, w, generator=g) * 3 for h, w in sizes] if task != 'bbox' else None
    metas = [dict(img_shape=(sizes[0][0] * 8 - 7 * i, sizes[0][1] * 8 - 11 * i, 3), scale_factor=np.array([1.25, 1.25, 1.25, 1.25], np.float32),
                  pad_shape=(sizes[0][0] * 8, sizes[0][1] * 8, 3), flip=False) for i in range(B)]
    return cls, box, lm, metas
