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
