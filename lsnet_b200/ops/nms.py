"""Inference-side post-processing with the reference's call signatures: ``nms`` / ``batched_nms``
(mmdet/ops/nms/nms_wrapper.py:7-157) and ``multiclass_nms_lsvr`` (mmdet/core/post_processing/bbox_nms.py:60-99) on
liblsnet_sm100's device-side greedy NMS (lsnet_nms: bitmask tiles + one-CTA sweep, no host round trip inside)."""
import ctypes

import torch

from .. import lib as L


def nms(dets, iou_thr):
    """dets (N,5) [x1,y1,x2,y2,score] on the device -> (dets[keep], keep) with keep in descending-score order."""
    n = dets.shape[0]
    if n == 0:
        return dets, dets.new_zeros(0, dtype=torch.long)
    order = dets[:, 4].sort(0, descending=True)[1]
    boxes = dets[order, :4].float().contiguous()
    ws_bytes = L.load().lsnet_nms_workspace_size(L.c_int(n))
    ws = torch.empty(int(ws_bytes), device=dets.device, dtype=torch.uint8)
    keep = torch.empty(n, device=dets.device, dtype=torch.int32)
    cnt = torch.zeros(1, device=dets.device, dtype=torch.int32)
    L.call('lsnet_nms', L.ptr(boxes), L.c_int(n), L.c_f(float(iou_thr)), L.ptr(ws), ctypes.c_size_t(ws_bytes),
           L.ptr(keep), L.ptr(cnt), L.stream())
    k = int(cnt.item())                       # the one synchronisation of the decode (the reference syncs here too)
    inds = order[keep[:k].long()]
    return dets[inds], inds


def batched_nms(bboxes, scores, inds, nms_cfg, class_agnostic=False):
    """nms_wrapper.py:119-157: per-class NMS through a per-class coordinate offset (kept in fp32 exactly as the reference
    does, so borderline IoU decisions are the same)."""
    cfg = dict(nms_cfg)
    class_agnostic = cfg.pop('class_agnostic', class_agnostic)
    if class_agnostic:
        boxes_for_nms = bboxes
    else:
        max_coordinate = bboxes.max()
        offsets = inds.to(bboxes) * (max_coordinate + 1)
        boxes_for_nms = bboxes + offsets[:, None]
    nms_type = cfg.pop('type', 'nms')
    if nms_type != 'nms':
        raise NotImplementedError(f'nms type {nms_type!r} (LSNet configs use plain nms)')
    dets, keep = nms(torch.cat([boxes_for_nms, scores[:, None]], -1), cfg.get('iou_thr', 0.5))
    return torch.cat([bboxes[keep], dets[:, -1:]], -1), keep


def multiclass_nms_lsvr(multi_bboxes, multi_pts, multi_scores, npts, score_thr, nms_cfg, max_num=-1, score_factors=None):
    """bbox_nms.py:60-99: every (box, class) pair above score_thr competes in a per-class NMS; the landmark vectors
    follow their boxes."""
    num_classes = multi_scores.size(1) - 1
    if multi_bboxes.shape[1] > 4:
        bboxes = multi_bboxes.view(multi_scores.size(0), -1, 4)
    else:
        bboxes = multi_bboxes[:, None].expand(-1, num_classes, 4)
    pts = multi_pts[:, None].expand(-1, num_classes, multi_pts.shape[-1])
    scores = multi_scores[:, :-1]
    valid_mask = scores > score_thr
    bboxes = bboxes[valid_mask]
    pts = pts[valid_mask]
    if score_factors is not None:
        scores = scores * score_factors[:, None]
    scores = scores[valid_mask]
    labels = valid_mask.nonzero()[:, 1]
    if bboxes.numel() == 0:
        return (multi_bboxes.new_zeros((0, 5)), pts.new_zeros((0, npts * 2)),
                multi_bboxes.new_zeros((0,), dtype=torch.long))
    dets, keep = batched_nms(bboxes, scores, labels, nms_cfg)
    if max_num > 0:
        dets, keep = dets[:max_num], keep[:max_num]
    return dets, pts[keep], labels[keep]
