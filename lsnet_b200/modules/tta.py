"""Multi-scale / flip testing by instance voting: the host-side half of ``LSDetector.aug_test_vote``
(mmdet/models/detectors/lsnet.py:138-409) and ``instance_mapping_back`` with its landmark flips
(mmdet/core/bbox/transforms.py:30-137).  The per-augmentation forward + decode + NMS run on the GPU through the same
kernels as ``simple_test``; what is here — a few hundred detections per image — is bookkeeping on the host, as in the
reference.
"""
import numpy as np
import torch

from ..datasets.transforms import KEYPOINT_FLIP_PAIRS


def remove_boxes(boxes, min_scale, max_scale):
    """lsnet.py:159-164: indices of boxes whose area lies in [min_scale², max_scale²]."""
    areas = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    return torch.nonzero((areas >= min_scale * min_scale) & (areas <= max_scale * max_scale), as_tuple=False).squeeze(1)


def flip_vectors(vectors, img_shape, task, direction='horizontal'):
    """Landmark vectors of a flipped image back into the un-flipped frame (transforms.py:30-88).  'bbox': the four
    extreme points (top, left, bottom, right) — mirror and swap left/right (or top/bottom); 'segm': mirror and
    re-order the ring as (p0, p_{n-1}, …, p_1) so it stays clockwise from the same start; pose: mirror and swap the
    left/right keypoint pairs."""
    assert direction in ('horizontal', 'vertical')
    v = vectors.clone()
    if task == 'bbox':
        if direction == 'horizontal':
            w = img_shape[1]
            v[..., 0::8], v[..., 4::8] = w - vectors[..., 0::8], w - vectors[..., 4::8]
            v[..., 2::8], v[..., 3::8] = w - vectors[..., 6::8], vectors[..., 7::8]
            v[..., 6::8], v[..., 7::8] = w - vectors[..., 2::8], vectors[..., 3::8]
        else:
            h = img_shape[0]
            v[..., 3::8], v[..., 7::8] = h - vectors[..., 3::8], h - vectors[..., 7::8]
            v[..., 0::8], v[..., 1::8] = vectors[..., 4::8], h - vectors[..., 5::8]
            v[..., 4::8], v[..., 5::8] = vectors[..., 0::8], h - vectors[..., 1::8]
        return v
    idx, dim = (0, img_shape[1]) if direction == 'horizontal' else (1, img_shape[0])
    if v.shape[0] == 0:
        if task == 'segm':
            v[:, idx::2] = dim - v[:, idx::2]
        return v
    v[:, idx::2] = dim - v[:, idx::2]
    p = v.reshape(v.shape[0], -1, 2)
    if task == 'segm':
        p = torch.cat([p[:, :1], torch.flip(p[:, 1:], [1])], 1)
    else:
        p = p.clone()
        for a, b in KEYPOINT_FLIP_PAIRS:
            p[:, [a, b]] = p[:, [b, a]]
    return p.reshape(v.shape[0], -1)


def instance_mapping_back(bboxes, vectors, img_shape, scale_factor, flip, task, flip_direction='horizontal'):
    """transforms.py:115-137: detections of one augmentation -> original image coordinates."""
    if flip:
        b = bboxes.clone()
        if flip_direction == 'vertical':
            b[..., 1::4], b[..., 3::4] = img_shape[0] - bboxes[..., 3::4], img_shape[0] - bboxes[..., 1::4]
        else:
            b[:, 0::4], b[:, 2::4] = img_shape[1] - bboxes[:, 2::4], img_shape[1] - bboxes[:, 0::4]
        bboxes, vectors = b, flip_vectors(vectors, img_shape, 'bbox' if task == 'bbox' else
                                         ('segm' if task == 'segm' else 'pose'), flip_direction)
    sf = np.asarray(scale_factor, np.float32).reshape(-1)
    sf = torch.as_tensor(np.repeat(sf, 4) if sf.size == 1 else sf, device=bboxes.device)     # a bare float: same in x and y
    return bboxes.view(-1, 4) / sf, vectors / sf[:2].repeat(vectors.shape[1] // 2)


def instances_vote(boxes, vectors, scores, vote_thresh=0.66):
    """lsnet.py:236-298: greedy clustering of one class's detections from all augmentations.  Repeatedly take the
    best-scoring remaining detection and everything overlapping it with IoU >= ``vote_thresh``; a cluster of one is
    kept as it is, a larger cluster becomes ONE detection (score-weighted mean of boxes and landmark vectors, the
    cluster's best score) plus its members re-scored by ``score · (1 − IoU)`` where that is still >= 0.05 (soft
    suppression; the seed itself has IoU 1 and disappears).  Fewer than two detections of the class: nothing is
    returned at all (the reference's ``<= 1`` guard).  Returns (boxes (n,4), vectors (n,V), scores (n,)) float32 on
    ``boxes.device`` sorted by score."""
    dev = boxes.device
    nv = vectors.shape[1]
    det = np.concatenate([boxes.detach().cpu().numpy(), scores.detach().cpu().numpy().reshape(-1, 1),
                          vectors.detach().cpu().numpy()], axis=1).astype(np.float64 if boxes.dtype == torch.float64
                                                                        else np.float32)
    if det.shape[0] <= 1:
        return boxes.new_zeros((0, 4)), vectors.new_zeros((0, nv)), scores.new_zeros((0,))
    det = det[det[:, 4].argsort()[::-1]]
    out = []
    while det.shape[0] > 0:
        area = (det[:, 2] - det[:, 0]) * (det[:, 3] - det[:, 1])
        w = np.maximum(0.0, np.minimum(det[0, 2], det[:, 2]) - np.maximum(det[0, 0], det[:, 0]))
        h = np.maximum(0.0, np.minimum(det[0, 3], det[:, 3]) - np.maximum(det[0, 1], det[:, 1]))
        inter = w * h
        iou = inter / np.maximum(area[0] + area - inter, 1e-6)
        iou[0] = 1
        member = iou >= vote_thresh
        group, giou = det[member], iou[member]
        det = det[~member]
        if group.shape[0] <= 1:
            out.append(group)
            continue
        soft = group.copy()
        soft[:, 4] = soft[:, 4] * (1 - giou)
        soft = soft[soft[:, 4] >= 0.05]
        sc = group[:, 4:5]
        merged = np.zeros((1, 5 + nv), det.dtype)
        merged[0, :4] = (group[:, :4] * sc).sum(0) / sc.sum()
        merged[0, 5:] = (group[:, 5:] * sc).sum(0) / sc.sum()
        merged[0, 4] = group[:, 4].max()
        out.append(merged)
        if soft.shape[0]:
            out.append(soft)
    dets = np.concatenate(out, 0)
    dets = dets[dets[:, 4].argsort()[::-1]]
    t = torch.from_numpy(np.ascontiguousarray(dets)).float().to(dev)
    return t[:, :4], t[:, 5:], t[:, 4]


def vote_merge(aug_dets, img_metas, task, num_classes, num_vectors, scale_ranges, max_per_img=1000):
    """lsnet.py:300-365 after the per-augmentation decode: scale-range filter (one range per scale; augmentations come
    in (plain, flipped) pairs, hence ``i // 2``), mapping back, per-class voting, top ``max_per_img`` by score.
    ``aug_dets``: [(det_bboxes (n,5), det_vectors (n,V), det_labels (n,))] per augmentation; ``img_metas``: the
    matching meta dicts.  Returns merged (det_bboxes (m,5), det_vectors (m,V), det_labels (m,)) in ORIGINAL image
    coordinates."""
    B, V, L = [], [], []
    for i, ((b, v, l), meta) in enumerate(zip(aug_dets, img_metas)):
        lo, hi = scale_ranges[i // 2]
        keep = remove_boxes(b, lo, hi)
        b, v, l = b[keep].clone(), v[keep], l[keep]
        b[:, :4], v = instance_mapping_back(b[:, :4], v, meta['img_shape'], meta['scale_factor'], meta['flip'], task,
                                            meta.get('flip_direction', 'horizontal'))
        B.append(b), V.append(v), L.append(l)
    B, V, L = torch.cat(B), torch.cat(V), torch.cat(L)
    ob, ov, ol = [], [], []
    for j in range(num_classes):
        idx = (L == j).nonzero().squeeze(1)
        bj, vj, sj = instances_vote(B[idx, :4].view(-1, 4), V[idx], B[idx, 4])
        if len(bj) > 0:
            ob.append(torch.cat([bj, sj[:, None]], 1))
            ov.append(vj)
            ol.append(torch.full((bj.shape[0],), j, dtype=torch.int64, device=sj.device))
    if not ob:
        return B.new_zeros((0, 5)), B.new_zeros((0, num_vectors * 2)), B.new_zeros((0,), dtype=torch.long)
    ob, ov, ol = torch.cat(ob), torch.cat(ov), torch.cat(ol)
    if ob.shape[0] > max_per_img:
        thr, _ = torch.kthvalue(ob[:, 4].cpu(), ob.shape[0] - max_per_img + 1)
        keep = torch.nonzero(ob[:, 4] >= thr.item(), as_tuple=False).squeeze(1)
        ob, ov, ol = ob[keep], ov[keep], ol[keep]
    return ob, ov, ol
