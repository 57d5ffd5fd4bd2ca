"""LSDetector: backbone -> neck -> LSHead.forward_train, with the reference's registry name, constructor arguments,
``forward`` / ``train_step`` / ``_parse_losses`` contract (mmdet/models/detectors/lsnet.py:14-56, single_stage.py:16-57,
base.py:161-243) and the single-scale inference path ``simple_test`` (lsnet.py:58-95).  Test-time augmentation /
instance voting (``aug_test_vote``, lsnet.py:97-409) is out of scope."""
from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.nn as nn

from ..registry import DETECTORS, build_backbone, build_head, build_neck


def parse_losses(losses, sync_log=False):
    """BaseDetector._parse_losses (mmdet/models/detectors/base.py:176-209).  The reference all-reduces and .item()s
    every logged scalar each step; here the scalars stay on the device and are reduced lazily in ONE collective when
    ``sync_log`` is set (logging interval)."""
    log_vars = OrderedDict()
    for name, value in losses.items():
        if isinstance(value, torch.Tensor):
            log_vars[name] = value.mean()
        elif isinstance(value, (list, tuple)):
            if all(v.dim() == 0 for v in value) and len(value) > 1:
                log_vars[name] = torch.stack(list(value)).sum()       # one reduction instead of len(value) scalar kernels
            else:
                log_vars[name] = sum(v.mean() for v in value)
        else:
            raise TypeError(f'{name} is not a tensor or list of tensors')
    loss = sum(v for k, v in log_vars.items() if 'loss' in k)
    log_vars['loss'] = loss
    if sync_log:
        flat = torch.stack([v.detach().float() for v in log_vars.values()])
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(flat.div_(dist.get_world_size()))
        log_vars = OrderedDict((k, v) for k, v in zip(log_vars.keys(), flat.tolist()))
    return loss, log_vars


@DETECTORS.register_module()
class LSDetector(nn.Module):

    def __init__(self, backbone, neck, bbox_head, train_cfg=None, test_cfg=None, pretrained=None):
        super().__init__()
        self.backbone = build_backbone(dict(backbone))
        self.neck = build_neck(dict(neck)) if neck is not None else None
        head_cfg = dict(bbox_head)
        head_cfg.update(train_cfg=train_cfg, test_cfg=test_cfg)
        self.bbox_head = build_head(head_cfg)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.init_weights(pretrained=pretrained)

    with_neck = property(lambda self: self.neck is not None)

    def init_weights(self, pretrained=None):       # single_stage.py:34-49
        self.backbone.init_weights(pretrained=pretrained)
        if self.with_neck:
            self.neck.init_weights()
        self.bbox_head.init_weights()

    def extract_feat(self, img):
        """single_stage.py:51-57.  The trunk runs in bf16 channels_last (cuDNN) under autocast."""
        img = img.contiguous(memory_format=torch.channels_last)
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=img.is_cuda):
            x = self.backbone(img)
        if self.with_neck:
            x = self.neck(x)
        return x

    def forward_train(self, img, img_metas, gt_bboxes, gt_labels, gt_masks=None, gt_extremes=None, gt_keypoints=None,
                      gt_bboxes_ignore=None):
        x = self.extract_feat(img)
        return self.bbox_head.forward_train(x, img_metas, gt_bboxes, gt_extremes, gt_keypoints, gt_masks, gt_labels,
                                            gt_bboxes_ignore)

    def forward(self, img, img_metas, return_loss=True, **kwargs):
        if return_loss:
            return self.forward_train(img, img_metas, **kwargs)
        return self.forward_test(img, img_metas, **kwargs)

    def forward_test(self, imgs, img_metas, **kwargs):
        """base.py:118-159: one augmentation = simple_test; several (MultiScaleFlipAug lists) = aug_test."""
        if torch.is_tensor(imgs):
            imgs, img_metas = [imgs], [img_metas]
        if len(imgs) != len(img_metas):
            raise ValueError(f'num of augmentations ({len(imgs)}) != num of image meta ({len(img_metas)})')
        if len(imgs) == 1:
            return self.simple_test(imgs[0], img_metas[0], **kwargs)
        if imgs[0].shape[0] != 1:
            raise ValueError('aug test does not support inference with batch size > 1 (base.py:150-151)')
        return self.aug_test(imgs, img_metas, **kwargs)

    def aug_test(self, imgs, img_metas, rescale=False, show=False, out_dir=False):
        """lsnet.py:402-409.  ``test_cfg.method == 'vote'`` is the multi-scale test of the reference's result tables;
        the plain box-merge variant ('simple', bbox task only, mmdet's generic merge_aug_results + multiclass_nms) is
        not built."""
        method = self.test_cfg.get('method', 'simple') if self.test_cfg is not None else 'simple'
        if method == 'simple':
            raise NotImplementedError("test_cfg.method='simple' (aug_test_simple, lsnet.py:102-136) is not built: use "
                                      "method='vote' with scale_ranges")
        return self.aug_test_vote(imgs, img_metas, rescale, show, out_dir)

    @torch.no_grad()
    def aug_test_vote(self, imgs, img_metas, rescale=False, show=False, out_dir=False):
        """lsnet.py:300-400: decode + NMS per augmentation on the device (one image each), then the scale-range filter,
        mapping back and per-class instance voting of ``modules/tta.py``; returns the per-class result lists of
        ``simple_test`` for ONE image."""
        import numpy as np
        from . import tta
        head = self.bbox_head
        dets, metas = [], []
        for img, meta in zip(imgs, img_metas):
            outs = head(self.extract_feat(img))
            dets.append(head.get_bboxes(*outs, meta, rescale=False, nms=True)[0])
            metas.append(meta[0])
        boxes, vecs, labels = tta.vote_merge(dets, metas, head.task, head.num_classes, head.num_vectors,
                                             self.test_cfg['scale_ranges'])
        if not rescale:      # back into the first augmentation's frame (lsnet.py:367-374)
            sf = np.asarray(metas[0]['scale_factor'], np.float32)
            boxes = boxes.clone()
            boxes[:, :4] *= boxes.new_tensor(sf)
            vecs = vecs * vecs.new_tensor(np.tile(sf[:2], vecs.shape[1] // 2))
        if head.task in ('pose_bbox', 'pose_kbox') and not (show or out_dir):
            big = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]) > 1024        # lsnet.py:389-396
            boxes, vecs, labels = boxes[big], vecs[big], labels[big]
        return self._to_results(boxes, vecs, labels)

    def _to_results(self, boxes, pts, labels):
        """``bbox_extreme2result`` / ``bbox_poly2result`` (mmdet/core/bbox/transforms.py:198-218): per-class numpy lists."""
        import numpy as np
        nc, nv = self.bbox_head.num_classes, self.bbox_head.num_vectors
        width = 8 if self.bbox_head.task == 'bbox' else 2 * nv
        if boxes.shape[0] == 0:
            return [[np.zeros((0, 5), np.float32) for _ in range(nc)], [np.zeros((0, width), np.float32) for _ in range(nc)]]
        b, p, l = boxes.cpu().numpy(), pts.cpu().numpy(), labels.cpu().numpy()
        return [[b[l == i] for i in range(nc)], [p[l == i] for i in range(nc)]]

    @torch.no_grad()
    def simple_test(self, img, img_metas, rescale=False, show=False, out_dir=False):
        """lsnet.py:58-95: decode + NMS per image, converted to the per-class numpy lists of ``bbox_extreme2result`` /
        ``bbox_poly2result`` (mmdet/core/bbox/transforms.py:198-218)."""
        x = self.extract_feat(img)
        outs = self.bbox_head(x)
        dets = self.bbox_head.get_bboxes(*outs, img_metas, rescale=rescale)
        return [self._to_results(boxes, pts, labels) for boxes, pts, labels in dets]

    def _parse_losses(self, losses, sync_log=False):
        return parse_losses(losses, sync_log)

    def train_step(self, data, optimizer=None, sync_log=False):
        """base.py:211-243."""
        losses = self(**data)
        loss, log_vars = self._parse_losses(losses, sync_log)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(data['img_metas']))
