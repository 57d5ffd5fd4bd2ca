"""CentroidAssigner / ATSSAssigner / PseudoSampler with the reference's names, constructor arguments and
``assign`` signatures (mmdet/core/bbox/assigners/centroid_assigner.py:26-93, atss_assigner.py:29-164,
samplers/pseudo_sampler.py:23-41), on the batched CUDA kernels.  The per-image ``assign`` API is kept for drop-in use;
LSHead calls the batched ops directly."""
import torch

from .. import ops
from ..registry import BBOX_ASSIGNERS, BBOX_SAMPLERS


class AssignResult:
    """mmdet/core/bbox/assigners/assign_result.py:43-49."""

    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts, self.gt_inds, self.max_overlaps, self.labels = num_gts, gt_inds, max_overlaps, labels


def _pyramid_from_points(points, device):
    """Rebuild the level geometry from the concatenated [x, y, stride] rows PointGenerator produced."""
    strides = points[:, 2]
    uniq = torch.unique_consecutive(strides).tolist()
    sizes = []
    for s in uniq:
        sel = points[strides == s]
        w = int(round(float(sel[:, 0].max()) / s)) + 1
        h = int(round(float(sel[:, 1].max()) / s)) + 1
        assert h * w == sel.shape[0], 'points must be full row-major PointGenerator grids'
        sizes.append((h, w))
    big = 1 << 30
    return ops.Pyramid(sizes, uniq, [(big, big)], device)


def _labels_of(gt_inds, gt_labels):
    if gt_labels is None:
        return None
    labels = gt_inds.new_full(gt_inds.shape, -1)
    pos = gt_inds > 0
    labels[pos] = gt_labels[gt_inds[pos] - 1]
    return labels


@BBOX_ASSIGNERS.register_module()
class CentroidAssigner:

    def __init__(self, scale=4, pos_num=3, iou_type='center'):
        if pos_num != 1 or iou_type != 'center':
            raise NotImplementedError("B200 CentroidAssigner covers the shipped configs: pos_num=1, iou_type='center'")
        self.scale, self.pos_num, self.iou_type = scale, pos_num, iou_type

    def assign(self, points, gt_bboxes, gt_extreme_pts=None, gt_bboxes_ignore=None, gt_labels=None):
        num_gts, num_points = gt_bboxes.shape[0], points.shape[0]
        if num_gts == 0 or num_points == 0:
            inds = points.new_full((num_points,), 0, dtype=torch.long)
            labels = None if gt_labels is None else points.new_full((num_points,), -1, dtype=torch.long)
            return AssignResult(num_gts, inds, None, labels=labels)
        pyr = _pyramid_from_points(points, points.device)
        a = ops.centroid_assign(pyr, gt_bboxes.float().contiguous()[None],
                                torch.tensor([num_gts], dtype=torch.int32, device=points.device), float(self.scale))
        inds = a[0].long() + 1
        return AssignResult(num_gts, inds, None, labels=_labels_of(inds, gt_labels))


@BBOX_ASSIGNERS.register_module()
class ATSSAssigner:

    def __init__(self, topk, iou_calculator=dict(type='BboxOverlaps2D')):
        self.topk = topk

    def assign(self, bboxes, num_level_bboxes, gt_bboxes, gt_bboxes_ignore=None, gt_labels=None):
        bboxes = bboxes[:, :4]
        num_gt, num_bboxes = gt_bboxes.size(0), bboxes.size(0)
        if num_gt == 0 or num_bboxes == 0:
            inds = bboxes.new_full((num_bboxes,), 0, dtype=torch.long)
            labels = None if gt_labels is None else bboxes.new_full((num_bboxes,), -1, dtype=torch.long)
            return AssignResult(num_gt, inds, bboxes.new_zeros((num_bboxes,)), labels=labels)
        big = 1 << 30
        pyr = ops.Pyramid([(1, int(n)) for n in num_level_bboxes], [2.0 ** (i + 3) for i in range(len(num_level_bboxes))],
                          [(big, big)], bboxes.device)
        a, mo = ops.atss_assign(pyr, bboxes.float().contiguous()[None], gt_bboxes.float().contiguous()[None],
                                torch.tensor([num_gt], dtype=torch.int32, device=bboxes.device), self.topk,
                                want_overlaps=True)
        inds = a[0].long() + 1
        return AssignResult(num_gt, inds, mo[0], labels=_labels_of(inds, gt_labels))


@BBOX_SAMPLERS.register_module()
class PseudoSampler:
    """pseudo_sampler.py:23-41: every assigned point is a positive, the rest negatives (no sampling)."""

    def __init__(self, **kwargs):
        pass

    def sample(self, assign_result, bboxes, gt_bboxes, **kwargs):
        pos = torch.nonzero(assign_result.gt_inds > 0, as_tuple=False).squeeze(-1).unique()
        neg = torch.nonzero(assign_result.gt_inds == 0, as_tuple=False).squeeze(-1).unique()
        return dict(pos_inds=pos, neg_inds=neg, pos_assigned_gt_inds=assign_result.gt_inds[pos] - 1)
