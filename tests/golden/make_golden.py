"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference through
oracle/ref_harness.py; its CUDA-only DCN/focal entry points backed by the C restatement) on the seeded inputs of
synth.py.  Only reference OUTPUTS are stored; inputs and weights are rebuilt from seeds by the tests.

    python tests/golden/make_golden.py        # needs /root/reference (this container only)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth  # noqa: E402
from oracle import init as oinit  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CFG = {'bbox': 'lsnet_bbox_r50_fpn_1x_coco.py', 'segm': 'lsnet_segm_r50_fpn_1x_coco.py',
       'pose_bbox': 'lsnet_pose_bbox_r50_fpn_1x_coco.py'}
GRAD_KEYS = ['backbone.layer2.0.conv1.weight', 'backbone.layer4.2.conv3.weight', 'neck.lateral_convs.0.conv.weight',
             'neck.fpn_convs.3.conv.weight', 'bbox_head.cls_convs.0.conv.weight',
             'bbox_head.cls_convs.2.conv.conv_offset.weight', 'bbox_head.pts_cls_conv.weight',
             'bbox_head.pts_cls_out.bias']


class _PM:   # stand-in for mmdet PolygonMasks: only .masks / .areas are read (lsnet_head.py:1717-1737)
    pass


def detector_golden():
    ns = rh.load()
    for task, cfgname in CFG.items():
        model, _ = rh.build_reference_detector(cfgname)
        sd = oinit.make_state_dict(task, seed=11)
        model.load_state_dict(sd)
        model.train()
        rec = {}
        for case, seed in enumerate((101, 202)):
            d = synth.detector_batch(task, seed)
            kw = {}
            if task == 'bbox':
                kw['gt_extremes'] = d['gt_extremes']
            if task == 'segm':
                gms = []
                for P in d['raw_polys']:
                    pm = _PM(); pm.areas = None
                    pm.masks = [[p.numpy().reshape(-1)] for p in P]
                    gms.append(pm)
                kw['gt_masks'] = gms
            if task == 'pose_bbox':
                kw['gt_keypoints'] = [k.clone() for k in d['gt_keypoints_vs']]
            gtb = d['src_bboxes'] if task == 'segm' else d['gt_bboxes']
            model.zero_grad()
            losses = model(img=d['img'], img_metas=d['img_metas'], gt_bboxes=gtb, gt_labels=d['gt_labels'],
                           return_loss=True, **kw)
            total = sum(sum(v) for v in losses.values())
            total.backward()
            for k, v in losses.items():
                rec[f'c{case}.{k}'] = np.array([float(x) for x in v], np.float64)
            params = dict(model.named_parameters())
            for k in GRAD_KEYS + [k for k in params if 'pose' in k or 'segm' in k][:4]:
                if k in params and params[k].grad is not None:
                    g = params[k].grad
                    rec[f'c{case}.gradnorm.{k}'] = np.array(float(g.norm()), np.float64)
                    rec[f'c{case}.gradhead.{k}'] = g.reshape(-1)[:16].numpy().astype(np.float32)
        np.savez_compressed(os.path.join(OUT, f'detector_{task}.npz'), **rec)
        print('detector', task, len(rec))


def loss_golden():
    ns = rh.load()
    from mmdet.models.losses.cross_iou_loss import CrossIOULoss
    from mmdet.models.losses.focal_loss import FocalLoss
    import oracle.lsnet_oracle as O
    rec = {}
    for lt in ('bbox', 'polygon', 'keypoint'):
        for seed in range(4):
            r = synth.loss_rows(lt, 500 + seed)
            tgt, sel = O.directional_targets(r['gt'], r['anchor'], r['weight'])   # inputs; checked vs reference below
            pred = r['pred'].clone().requires_grad_(True)
            mod = CrossIOULoss(loss_type=lt, loss_weight=1.5)
            kw = dict(anchor_pts=r['anchor'], bbox_gt=None if lt == 'keypoint' else r['bbox_gt'], pos_inds=sel)
            if lt == 'keypoint':
                kw['vs'] = r['vs'].clone()
            loss = mod(pred, tgt.clone(), r['weight'], avg_factor=7.0, **kw)
            loss.backward()
            rec[f'{lt}.{seed}.loss'] = np.array(float(loss), np.float64)
            rec[f'{lt}.{seed}.grad'] = pred.grad.numpy()
    # focal
    rng = np.random.RandomState(9)
    logits = torch.from_numpy((rng.randn(200, 80) * 3).astype(np.float32)).requires_grad_(True)
    labels = torch.from_numpy(rng.randint(0, 81, 200))
    w = torch.from_numpy((rng.rand(200) > 0.1).astype(np.float32))
    fl = FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0)
    loss = fl(logits, labels, w, avg_factor=13.0)
    loss.backward()
    rec['focal.loss'] = np.array(float(loss), np.float64)
    rec['focal.grad'] = logits.grad.numpy()
    # directional targets straight from the reference head (get_bbox_gt_reg / get_poly_gt_reg)
    model, _ = rh.build_reference_detector(CFG['bbox'])
    head = model.bbox_head
    for lt, NP in (('bbox', 5), ('polygon', 37)):
        r = synth.loss_rows(lt, 777)
        anc3 = torch.cat([r['anchor'], torch.ones(len(r['anchor']), 1)], 1)
        if lt == 'bbox':
            t, s = head.get_bbox_gt_reg(r['gt'], anc3, r['weight'])
        else:
            head.num_vectors = 36
            t, s = head.get_poly_gt_reg(r['gt'], anc3, r['weight'])
            head.num_vectors = 4
        rec[f'dirtgt.{lt}.t'] = t.numpy()
        rec[f'dirtgt.{lt}.s'] = s.numpy()
    np.savez_compressed(os.path.join(OUT, 'losses.npz'), **rec)
    print('losses', len(rec))


def assign_golden():
    ns = rh.load()
    from mmdet.core.bbox.assigners import ATSSAssigner, CentroidAssigner
    rec = {}
    ca, aa = CentroidAssigner(scale=4, pos_num=1, iou_type='center'), ATSSAssigner(topk=9)
    for seed in range(6):
        c = synth.assign_case(900 + seed, G=1 + 3 * seed)
        r1 = ca.assign(c['points'], c['gt'], None, None, torch.arange(len(c['gt'])))
        r2 = aa.assign(c['pred'], c['num_level'], c['gt'], None, torch.arange(len(c['gt'])))
        rec[f'{seed}.centroid'] = r1.gt_inds.numpy()
        rec[f'{seed}.atss'] = r2.gt_inds.numpy()
        rec[f'{seed}.atss_max'] = r2.max_overlaps.numpy()
    np.savez_compressed(os.path.join(OUT, 'assign.npz'), **rec)
    print('assign', len(rec))


def decode_golden():
    """LSHead.get_bboxes of the reference (lsnet_head.py:1439-1668 + multiclass_nms_lsvr) on seeded head outputs; the
    reference's compiled nms_ext is absent here, its CPU entry point is served by oracle.decode_ref.greedy_nms."""
    from oracle import decode_ref as DR
    rh.load()
    sys.modules['mmdet.ops.nms.nms_ext'].nms = lambda dets, thr: DR.greedy_nms(dets, thr)
    rec = {}
    for task, cfgname in CFG.items():
        model, cfg = rh.build_reference_detector(cfgname)
        head = model.bbox_head
        none = [None] * 5
        for seed in (7, 8):
            cls, box, lm, metas = DR.synth_head_outputs(task, seed)
            args = dict(bbox=(cls, none, box, none, none, none, none), segm=(cls, none, none, none, lm, none, none),
                        pose_bbox=(cls, none, box, none, none, none, lm))[task]
            for rescale in (False, True):
                res = head.get_bboxes(*args, metas, rescale=rescale)
                for i, (b, p, l) in enumerate(res):
                    key = f'{task}.{seed}.{int(rescale)}.{i}'
                    rec[key + '.boxes'] = b.detach().numpy().astype(np.float32)
                    rec[key + '.pts'] = p.detach().numpy().astype(np.float32)
                    rec[key + '.labels'] = l.detach().numpy().astype(np.int64)
    np.savez_compressed(os.path.join(OUT, 'decode.npz'), **rec)
    print('decode', len(rec), {k: v.shape for k, v in list(rec.items())[:3]})


if __name__ == '__main__':
    torch.manual_seed(0)
    if 'decode' in sys.argv:
        decode_golden()
        sys.exit(0)
    loss_golden()
    assign_golden()
    detector_golden()
    decode_golden()
