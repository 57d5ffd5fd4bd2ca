"""Cross-IOU / focal loss autograd ops on liblsnet_sm100.so.  Reductions to the scalar loss are done with torch ops
on the tiny per-row / per-block outputs; avg_factor may be a device tensor (no host sync anywhere)."""
import torch
from torch.autograd import Function

from .. import lib as L

LOSS_TYPES = {'bbox': 0, 'polygon': 1, 'keypoint': 2}


def _f32c(t):
    return None if t is None else t.detach().float().contiguous()


def _scale_tensor(g, device):
    return (g.detach().float() if torch.is_tensor(g) else torch.tensor(float(g))).reshape(1).to(device).contiguous()


class _CrossIouRows(Function):
    """row_loss[n] = weight_row[n] * cross_iou(pred[n], target[n])  (dense form, cross_iou_loss.py:61-132)."""

    @staticmethod
    def forward(ctx, pred, target, pos_inds, weight_row, anchor_pts, bbox_gt, vs, loss_type, eps, alpha, stride):
        p = pred.detach().float().contiguous()
        N, D = p.shape
        t, s = _f32c(target), pos_inds.detach().to(torch.uint8).contiguous()
        w, a, b, v = _f32c(weight_row), _f32c(anchor_pts), _f32c(bbox_gt), _f32c(vs)
        Lk = v.shape[1] if v is not None else 0
        rows = torch.empty(N, device=p.device, dtype=torch.float32)
        L.call('lsnet_cross_iou_fwd', L.c_int(loss_type), L.ptr(p), L.ptr(t), L.ptr(s), L.ptr(w), L.ptr(a), L.ptr(b),
               L.ptr(v), L.c_int(N), L.c_int(D), L.c_int(Lk), L.c_f(eps), L.c_f(alpha), L.c_int(stride), L.ptr(rows),
               L.stream())
        ctx.saved = (p, t, s, w, a, b, v, loss_type, eps, alpha, stride, Lk)
        return rows

    @staticmethod
    def backward(ctx, grows):
        p, t, s, w, a, b, v, loss_type, eps, alpha, stride, Lk = ctx.saved
        N, D = p.shape
        # rows are reduced by sum / avg_factor upstream, so grows is one value broadcast over rows; keep generality
        # by folding the per-row upstream gradient into the row weights.
        wr = (w if w is not None else torch.ones(N, device=p.device)) * grows.float()
        one = torch.ones(1, device=p.device)
        dpred = torch.empty_like(p)
        # weight > 0 gates the row in the kernel: use |w*g| for the gate and restore the sign afterwards
        sign = torch.sign(wr)
        L.call('lsnet_cross_iou_bwd', L.c_int(loss_type), L.ptr(p), L.ptr(t), L.ptr(s), L.ptr(wr.abs().contiguous()),
               L.ptr(a), L.ptr(b), L.ptr(v), L.c_int(N), L.c_int(D), L.c_int(Lk), L.c_f(eps), L.c_f(alpha),
               L.c_int(stride), L.ptr(one), L.ptr(dpred), L.stream())
        dpred = dpred * sign.view(-1, 1)
        return dpred, None, None, None, None, None, None, None, None, None, None


def cross_iou_loss_rows(pred, target, pos_inds, weight_row=None, anchor_pts=None, bbox_gt=None, vs=None,
                        loss_type='bbox', eps=1e-6, alpha=0.2, stride=9):
    return _CrossIouRows.apply(pred, target, pos_inds, weight_row, anchor_pts, bbox_gt, vs, LOSS_TYPES[loss_type],
                               float(eps), float(alpha), int(stride))


class _CrossIouLevel(Function):
    """Fused per-level loss: sum over rows of the level's NHWC prediction map (LSHead.loss_single,
    lsnet_head.py:1064-1102) -> scalar sum (un-normalised)."""

    @staticmethod
    def forward(ctx, pred, assign, level_off, stride, base_scale, gt_pts, gt_bbox, gt_vs, loss_type, eps, alpha,
                pstride):
        # pred: (B, D, H, W) fp32, pixel-major memory
        B, D, H, W = pred.shape
        ldp = pred.stride(3)
        assert pred.dtype == torch.float32 and pred.stride(1) == 1 and pred.stride(2) == W * ldp
        NP = gt_pts.shape[2] // 2
        Lk = gt_vs.shape[2] if gt_vs is not None else 0
        rows = torch.empty(B * H * W, device=pred.device, dtype=torch.float32)
        args = (L.c_int(loss_type),)
        common = (L.ptr(pred), L.c_ll(ldp), L.c_int(D), L.ptr(assign), L.c_ll(assign.stride(0)), L.c_int(level_off),
                  L.c_int(B), L.c_int(H), L.c_int(W), L.c_f(stride), L.c_f(base_scale), L.ptr(gt_pts), L.ptr(gt_bbox),
                  L.ptr(gt_vs), L.c_int(gt_pts.shape[1]), L.c_int(NP), L.c_int(Lk), L.c_f(eps), L.c_f(alpha),
                  L.c_int(pstride))
        L.call('lsnet_cross_iou_level', *args, L.c_int(0), *common, L.ptr(rows), L.c_vp(0), L.c_vp(0), L.stream())
        ctx.saved = (pred, assign, gt_pts, gt_bbox, gt_vs, args, common)
        return rows.sum()

    @staticmethod
    def backward(ctx, g):
        pred, assign, gt_pts, gt_bbox, gt_vs, args, common = ctx.saved
        B, D, H, W = pred.shape
        scale = _scale_tensor(g, pred.device)
        dbuf = torch.empty((B, H, W, pred.stride(3)), device=pred.device, dtype=torch.float32)
        L.call('lsnet_cross_iou_level', *args, L.c_int(1), *common, L.c_vp(0), L.ptr(scale), L.ptr(dbuf), L.stream())
        return (dbuf.permute(0, 3, 1, 2)[:, :D],) + (None,) * 11


def cross_iou_level_loss(pred, assign, level_off, stride, base_scale, gt_pts, gt_bbox, gt_vs=None, loss_type='bbox',
                         eps=1e-6, alpha=0.2, pstride=9):
    return _CrossIouLevel.apply(pred, assign, int(level_off), float(stride), float(base_scale), gt_pts, gt_bbox, gt_vs,
                                LOSS_TYPES[loss_type], float(eps), float(alpha), int(pstride))


class _FocalSum(Function):
    """sum_n weight[n] * sum_c focal(logits[n,c], labels[n])  (sigmoid_focal_loss_cuda.cu:23-97)."""

    @staticmethod
    def forward(ctx, logits, labels, weight, gamma, alpha):
        # logits: (N, C) fp32 rows with pitch stride(0)
        assert logits.dtype == torch.float32 and logits.stride(1) == 1
        N, C = logits.shape
        labels = labels.to(torch.int32).contiguous()
        weight = _f32c(weight)
        nb = L.load().lsnet_focal_partial_count(L.c_ll(N), L.c_int(C))
        partial = torch.empty(nb, device=logits.device, dtype=torch.float32)
        L.call('lsnet_focal_fwd', L.ptr(logits), L.c_ll(logits.stride(0)), L.ptr(labels), L.ptr(weight), L.c_ll(N),
               L.c_int(C), L.c_f(gamma), L.c_f(alpha), L.ptr(partial), L.stream())
        ctx.saved = (logits, labels, weight, gamma, alpha)
        return partial.sum()

    @staticmethod
    def backward(ctx, g):
        logits, labels, weight, gamma, alpha = ctx.saved
        N, C = logits.shape
        scale = _scale_tensor(g, logits.device)
        d = torch.empty((N, logits.stride(0)), device=logits.device, dtype=torch.float32)
        if logits.stride(0) != C:
            d.zero_()
        L.call('lsnet_focal_bwd', L.ptr(logits), L.c_ll(logits.stride(0)), L.ptr(labels), L.ptr(weight), L.c_ll(N),
               L.c_int(C), L.c_f(gamma), L.c_f(alpha), L.ptr(scale), L.ptr(d), L.c_ll(d.stride(0)), L.stream())
        return d[:, :C], None, None, None, None


def sigmoid_focal_loss_sum(logits, labels, weight=None, gamma=2.0, alpha=0.25):
    return _FocalSum.apply(logits, labels, weight, float(gamma), float(alpha))


def directional_targets(gt_rows, anchor_pts, weight_row):
    """LSHead.get_bbox_gt_reg / get_poly_gt_reg (lsnet_head.py:402-454): (target [N,4NP] fp32, pos_inds [N,4NP] bool)."""
    g, a, w = _f32c(gt_rows), _f32c(anchor_pts[:, :2]), _f32c(weight_row)
    N, NP = g.shape[0], g.shape[1] // 2
    t = torch.empty((N, 4 * NP), device=g.device, dtype=torch.float32)
    s = torch.empty((N, 4 * NP), device=g.device, dtype=torch.uint8)
    L.call('lsnet_directional_targets', L.ptr(g), L.ptr(a), L.ptr(w), L.c_int(N), L.c_int(NP), L.ptr(t), L.ptr(s),
           L.stream())
    return t, s.bool()
