"""Inference decode on the device: lsnet_nms against the oracle's greedy NMS, LSHead.get_bboxes against
the oracle decode (itself pinned to the reference's get_bboxes by tests/test_oracle_decode.py) on identical head outputs,
and LSDetector.simple_test end to end."""
import numpy as np
import pytest
import torch

from oracle import decode_ref as DR
from test_oracle_decode import NV, STRIDES, TEST_CFG

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n,seed', [(1, 0), (63, 1), (64, 2), (65, 3), (1000, 4), (5000, 5)])
def test_nms_kernel_vs_oracle(n, seed):
    import lsnet_b200.ops as ops
    g = torch.Generator().manual_seed(seed)
    ctr = torch.rand(n, 2, generator=g) * 300
    wh = torch.rand(n, 2, generator=g) * 60 + 1
    dets = torch.cat([ctr - wh / 2, ctr + wh / 2, torch.rand(n, 1, generator=g)], 1)
    dets[::17, 2] = dets[::17, 0]                    # degenerate (zero-area) boxes: IoU 0/0 never suppresses
    keep_ref = DR.greedy_nms(dets, 0.6)
    out, keep = ops.nms(dets.cuda(), 0.6)
    assert keep.cpu().tolist() == keep_ref.tolist()
    assert torch.equal(out.cpu(), dets[keep_ref])


@pytest.mark.parametrize('task', ['bbox', 'segm', 'pose_bbox'])
def test_get_bboxes_vs_oracle(task):
    import lsnet_b200 as L
    from lsnet_b200.data import MODEL_CFG
    name = {'bbox': 'bbox_r50', 'segm': 'segm_r50', 'pose_bbox': 'pose_x101dcn'}[task]
    head_cfg = dict(MODEL_CFG[name]['model']['bbox_head'])
    head_cfg.update(train_cfg=None, test_cfg=TEST_CFG[task])
    head = L.build_head(head_cfg).cuda().eval()
    none = [None] * 5
    for seed in (7, 9):
        cls, box, lm, metas = DR.synth_head_outputs(task, seed)
        dev = lambda ts: None if ts is None else [t.cuda() for t in ts]
        args = dict(bbox=(dev(cls), none, dev(box), none, none, none, none),
                    segm=(dev(cls), none, none, none, dev(lm), none, none),
                    pose_bbox=(dev(cls), none, dev(box), none, none, none, dev(lm)))[task]
        for rescale in (False, True):
            got = head.get_bboxes(*args, metas, rescale=rescale)
            ref = DR.get_bboxes(task, NV[task], STRIDES, cls, box, lm, metas, TEST_CFG[task], rescale=rescale)
            for (b, p, l), (rb, rp, rl) in zip(got, ref):
                assert rb.shape[0] > 0
                assert torch.equal(l.cpu(), rl), (task, seed)
                # same fp32 arithmetic on both sides; sigmoid may differ in the last bit between CPU and GPU libm
                assert torch.allclose(b.cpu(), rb, rtol=1e-6, atol=1e-5)
                assert torch.allclose(p.cpu(), rp, rtol=1e-6, atol=1e-5)


def test_simple_test_end_to_end():
    import lsnet_b200 as L
    from lsnet_b200.data import MODEL_CFG, synthetic_batch
    cfg = MODEL_CFG['bbox_r50']
    torch.manual_seed(0)
    model = L.build_detector(cfg['model'], train_cfg=cfg['train_cfg'], test_cfg=TEST_CFG['bbox']).cuda().eval()
    torch.nn.init.constant_(model.bbox_head.pts_cls_out.bias, -1.0)          # scores above the 0.05 threshold
    b = synthetic_batch(0, batch=2, img_hw=(256, 320))
    res = model(img=b['img'].cuda(), img_metas=b['img_metas'], return_loss=False)
    assert len(res) == 2
    for boxes, pts in res:
        assert len(boxes) == 80 and len(pts) == 80
        n = sum(len(x) for x in boxes)
        assert 0 < n <= 100
        allb = np.concatenate(boxes)
        assert allb.shape[1] == 5 and np.isfinite(allb).all()
        assert (allb[:, 0] >= 0).all() and (allb[:, 2] <= 320).all() and (allb[:, 3] <= 256).all()
        assert np.concatenate(pts).shape[1] == 8


def _vote_model():
    import lsnet_b200 as L
    from lsnet_b200.data import MODEL_CFG
    cfg = MODEL_CFG['bbox_r50']
    torch.manual_seed(0)
    test_cfg = dict(TEST_CFG['bbox'], method='vote', scale_ranges=[[0, 100000], [0, 100000]])
    model = L.build_detector(cfg['model'], train_cfg=cfg['train_cfg'], test_cfg=test_cfg).cuda().eval()
    torch.nn.init.constant_(model.bbox_head.pts_cls_out.bias, -1.0)          # scores above the 0.05 threshold
    return model


def test_aug_test_vote_of_a_repeated_augmentation_is_simple_test():
    """Voting over the SAME augmentation twice gives back simple_test's detections: every proper box meets its copy
    (IoU 1), the pair merges into itself with its own score and the soft-suppressed copy scores 0; NMS has already
    separated distinct detections below the vote threshold.  A randomly initialised head also decodes inverted /
    empty boxes (x2 < x1 or y2 < y1): the scale filter drops those of negative area and the rest never overlap
    anything, as in the reference (lsnet.py:159-164, 236-262) -- so the claim is checked on the proper boxes, and in
    the other direction every voted detection must be one of simple_test's.  Matching is by nearest box with a
    budget of two misses each way: the forward is not bit-reproducible, a detection on the top-100 cut may come or go."""
    from lsnet_b200.data import synthetic_batch
    model = _vote_model()
    b = synthetic_batch(0, batch=1, img_hw=(256, 320))
    img = b['img'].cuda()
    meta = [dict(b['img_metas'][0], scale_factor=np.ones(4, np.float32), flip=False)]
    ref = model.simple_test(img, meta)[0]
    got = model(img=[img, img], img_metas=[meta, meta], return_loss=False)
    assert len(got) == 2 and len(got[0]) == 80 and len(got[1]) == 80
    miss_ref = miss_got = proper = 0
    for c in range(80):
        rb, rp, gb, gp = ref[0][c], ref[1][c], got[0][c], got[1][c]
        assert gb.shape[1] == 5 and gp.shape[1] == 8 and len(gb) == len(gp)

        def found(box, pts, pool_b, pool_p):
            if len(pool_b) == 0:
                return False
            d = np.abs(pool_b[:, :4] - box[:4]).max(1)
            j = int(d.argmin())
            return d[j] < 1e-2 and abs(pool_b[j, 4] - box[4]) < 1e-3 and np.abs(pool_p[j] - pts).max() < 1e-2
        for k in range(len(rb)):
            if rb[k, 2] - rb[k, 0] > 0.5 and rb[k, 3] - rb[k, 1] > 0.5:
                proper += 1
                miss_ref += not found(rb[k], rp[k], gb, gp)
        for k in range(len(gb)):
            miss_got += not found(gb[k], gp[k], rb, rp)
    assert proper >= 1 and miss_ref <= 2 and miss_got <= 2, (proper, miss_ref, miss_got)


def test_multi_scale_flip_pipeline_to_voted_result(tmp_path):
    """Image file -> MultiScaleFlipAug test pipeline (2 scales x flip) -> collate -> forward_test -> aug_test_vote."""
    import cv2
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import synth_coco as S
    from lsnet_b200 import datasets as D
    from lsnet_b200.registry import DATASETS
    for i in range(len(S.SIZES)):
        cv2.imwrite(str(tmp_path / f'img_{i}.png'), S.image(i))
    pipe = S.eval_pipeline(True)
    pipe[0]['img_scale'] = [(448, 256), (640, 384)]
    ds = DATASETS.get('CocoDataset')(ann_file=S.coco_dict(False), pipeline=[dict(type='LoadImageFromFile')] + pipe,
                                     img_prefix=str(tmp_path), test_mode=True)
    batch = D.collate([ds[0]])
    assert len(batch['img']) == 4 and [m[0]['flip'] for m in batch['img_metas']] == [False, True, False, True]
    model = _vote_model()
    res = model(img=[im.cuda() for im in batch['img']], img_metas=batch['img_metas'], return_loss=False, rescale=True)
    boxes, pts = res
    assert len(boxes) == 80 and len(pts) == 80
    allb = np.concatenate(boxes)
    h, w = S.SIZES[0]
    assert allb.shape[1] == 5 and np.isfinite(allb).all() and len(allb) > 0
    # rescale=True: original-image coordinates (decoded boxes are clipped to each augmentation's image before mapping back)
    assert allb[:, 0].min() >= -1e-3 and allb[:, 2].max() <= w + 1e-3 and allb[:, 3].max() <= h + 1e-3
