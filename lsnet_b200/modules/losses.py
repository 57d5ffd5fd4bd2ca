"""CrossIOULoss / FocalLoss with the reference's constructor arguments and forward signatures
(mmdet/models/losses/cross_iou_loss.py:134-172, focal_loss.py:119-186), computed by the CUDA kernels."""
import torch
import torch.nn as nn

from .. import ops
from ..registry import LOSSES


def _reduce(rows_sum_or_rows, reduction, avg_factor, n):
    """weight_reduce_loss semantics (mmdet/models/losses/utils.py:26-52) for an already weighted row vector."""
    loss = rows_sum_or_rows
    if avg_factor is None:
        if reduction == 'mean':
            return loss.sum() / n
        if reduction == 'sum':
            return loss.sum()
        return loss
    if reduction == 'mean':
        return loss.sum() / avg_factor
    if reduction == 'none':
        return loss
    raise ValueError('avg_factor can not be used with reduction="sum"')


@LOSSES.register_module()
class CrossIOULoss(nn.Module):

    def __init__(self, eps=1e-6, reduction='mean', loss_weight=1.0, loss_type='bbox', alpha=0.2, stride=9):
        super().__init__()
        self.eps, self.reduction, self.loss_weight = eps, reduction, loss_weight
        self.loss_type, self.alpha, self.stride = loss_type, alpha, stride

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, anchor_pts=None,
                bbox_gt=None, pos_inds=None, vs=None, **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        wrow = None
        if weight is not None:
            wrow = weight.mean(-1) if weight.dim() > 1 else weight
        # rows with zero weight contribute exactly 0 to value and gradient, which also covers the reference's
        # "no positive" early return (cross_iou_loss.py:153-154) without a host sync
        rows = ops.cross_iou_loss_rows(pred, target, pos_inds, wrow, None if anchor_pts is None else anchor_pts[:, :2],
                                       bbox_gt, vs, loss_type=self.loss_type, eps=self.eps, alpha=self.alpha,
                                       stride=self.stride)
        return self.loss_weight * _reduce(rows, reduction, avg_factor, pred.shape[0])


@LOSSES.register_module()
class FocalLoss(nn.Module):

    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction='mean', loss_weight=1.0):
        super().__init__()
        assert use_sigmoid is True, 'Only sigmoid focal loss supported now.'
        self.use_sigmoid, self.gamma, self.alpha = use_sigmoid, gamma, alpha
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        if reduction == 'none':
            raise NotImplementedError('FocalLoss(reduction="none") is not on the LSNet path')
        pred = pred.float()
        if pred.stride(1) != 1:
            pred = pred.contiguous()
        total = ops.sigmoid_focal_loss_sum(pred, target, weight, self.gamma, self.alpha)
        if avg_factor is None:
            total = total / pred.numel() if reduction == 'mean' else total
        else:
            total = total / avg_factor
        return self.loss_weight * total
