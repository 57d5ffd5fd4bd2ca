"""Landmark-target assignment ops (batched over images) on liblsnet_sm100.so."""
import numpy as np
import torch

from .. import lib as L


class Pyramid:
    """Host description of the point pyramid (PointGenerator grids, core/anchor/point_generator.py:17-37) plus the
    per-image valid extents derived from pad_shape (LSHead.get_points, lsnet_head.py:781-792)."""

    def __init__(self, sizes, strides, pad_shapes, device, valid_hw=None):
        self.sizes = [(int(h), int(w)) for h, w in sizes]
        self.strides = [float(s) for s in strides]
        self.n = len(self.sizes)
        self.num_level = [h * w for h, w in self.sizes]
        self.offsets = [0] + list(np.cumsum(self.num_level)[:-1])
        self.total = int(sum(self.num_level))
        self.h_arr = L.host_int_array([h for h, _ in self.sizes])
        self.w_arr = L.host_int_array([w for _, w in self.sizes])
        self.s_arr = L.host_float_array(self.strides)
        if valid_hw is not None:           # already on the device (static buffer of a captured step)
            self.valid_hw, self.B = valid_hw, valid_hw.shape[0]
            return
        valid = []
        for ph, pw in pad_shapes:
            valid.append([[min(int(np.ceil(ph / s)), h), min(int(np.ceil(pw / s)), w)]
                          for (h, w), s in zip(self.sizes, self.strides)])
        self.B = len(valid)
        self.valid_hw = torch.tensor(valid, dtype=torch.int32).to(device, non_blocking=True)

    def args(self):
        return (L.c_int(self.n), self.h_arr, self.w_arr, self.s_arr, L.ptr(self.valid_hw))


def centroid_assign(pyr, gt_bbox, gt_count, scale=4.0):
    """CentroidAssigner (pos_num=1, 'center').  gt_bbox [B,Gmax,4] fp32, gt_count [B] int32 -> assign [B,total] int32
    (-1 background, else GT index)."""
    B, Gmax = gt_bbox.shape[:2]
    dev = gt_bbox.device
    assign = torch.empty((B, pyr.total), device=dev, dtype=torch.int32)
    ws_d = torch.empty((B, max(Gmax, 1)), device=dev, dtype=torch.float32)
    ws_i = torch.empty((B, max(Gmax, 1)), device=dev, dtype=torch.int32)
    L.call('lsnet_centroid_assign', *pyr.args(), L.ptr(gt_bbox), L.ptr(gt_count), L.c_int(B), L.c_int(Gmax),
           L.c_f(scale), L.ptr(ws_d), L.ptr(ws_i), L.ptr(assign), L.stream())
    return assign


def atss_assign(pyr, boxes, gt_bbox, gt_count, topk=9, want_overlaps=False):
    """ATSSAssigner.  boxes [B,total,4] fp32 predicted init boxes."""
    B, Gmax = gt_bbox.shape[:2]
    dev = gt_bbox.device
    assign = torch.empty((B, pyr.total), device=dev, dtype=torch.int32)
    keys = torch.empty((B, pyr.total), device=dev, dtype=torch.int64)
    mo = torch.empty((B, pyr.total), device=dev, dtype=torch.float32) if want_overlaps else None
    L.call('lsnet_atss_assign', *pyr.args(), L.ptr(boxes), L.ptr(gt_bbox), L.ptr(gt_count), L.c_int(B), L.c_int(Gmax),
           L.c_int(topk), L.ptr(keys), L.ptr(assign), L.ptr(mo), L.stream())
    return (assign, mo) if want_overlaps else assign


def assign_targets(pyr, assign, gt_labels, num_classes):
    """labels int32 [B,total] (background = num_classes), label_weights fp32, num_pos int32 [B]."""
    B = assign.shape[0]
    dev = assign.device
    Gmax = gt_labels.shape[1] if gt_labels is not None else 0
    labels = torch.empty((B, pyr.total), device=dev, dtype=torch.int32)
    lw = torch.empty((B, pyr.total), device=dev, dtype=torch.float32)
    npos = torch.empty(B, device=dev, dtype=torch.int32)
    L.call('lsnet_assign_targets', *pyr.args(), L.ptr(assign), L.ptr(gt_labels), L.c_int(B), L.c_int(Gmax),
           L.c_int(num_classes), L.ptr(labels), L.ptr(lw), L.ptr(npos), L.stream())
    return labels, lw, npos


def pred_boxes(pyr, preds, polygon=False):
    """Predicted init boxes [B,total,4] from the per-level softplus'd maps (B, 4*NP, H_l, W_l) fp32 pixel-major."""
    B = preds[0].shape[0]
    boxes = torch.empty((B, pyr.total, 4), device=preds[0].device, dtype=torch.float32)
    for l, p in enumerate(preds):
        _, D, H, W = p.shape
        assert p.dtype == torch.float32 and p.stride(1) == 1 and p.stride(2) == W * p.stride(3)
        L.call('lsnet_pred_boxes', L.ptr(p), L.c_ll(p.stride(3)), L.c_int(D // 4), L.c_int(int(polygon)), L.c_int(B),
               L.c_int(H), L.c_int(W), L.c_f(pyr.strides[l]), L.c_int(int(pyr.offsets[l])), L.c_int(pyr.total),
               L.ptr(boxes), L.stream())
    return boxes
