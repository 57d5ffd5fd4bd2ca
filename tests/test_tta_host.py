"""Multi-scale / flip testing by voting (host half of LSDetector.aug_test_vote) against outputs of the
reference's own remove_boxes / merge_aug_vote_results / instances_vote on the synthetic per-augmentation detections of
tests/golden/synth_coco.py::tta_case (recorded in tests/golden/datapath.npz by make_golden_data.py)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import synth_coco as S  # noqa: E402

from lsnet_b200.modules import tta  # noqa: E402

G = np.load(os.path.join(HERE, 'golden', 'datapath.npz'))


def _case(task):
    dets, metas = S.tta_case(task)
    return [(torch.from_numpy(b), torch.from_numpy(v), torch.from_numpy(l)) for b, v, l in dets], metas


@pytest.mark.parametrize('task', ['bbox', 'segm', 'pose_bbox'])
def test_scale_filter_and_mapping_back(task):
    dets, metas = _case(task)
    B, V = [], []
    for i, ((b, v, l), m) in enumerate(zip(dets, metas)):
        keep = tta.remove_boxes(b, *S.TTA_SCALE_RANGES[i // 2])
        assert np.array_equal(keep.numpy(), G[f'tta_{task}_keep_{i}'])
        nb, nv = tta.instance_mapping_back(b[keep, :4], v[keep], m['img_shape'], m['scale_factor'], m['flip'], task)
        B.append(torch.cat([nb, b[keep, 4:]], 1))
        V.append(nv)
    assert np.array_equal(torch.cat(B).numpy(), G[f'tta_{task}_mapped_boxes'])
    assert np.array_equal(torch.cat(V).numpy(), G[f'tta_{task}_mapped_vectors'])
    assert any(m['flip'] for m in metas)


@pytest.mark.parametrize('task', ['bbox', 'segm', 'pose_bbox'])
def test_vote_merge_matches_reference(task):
    dets, metas = _case(task)
    nv = dets[0][1].shape[1] // 2
    boxes, vecs, labels = tta.vote_merge(dets, metas, task, 3, nv, S.TTA_SCALE_RANGES)
    assert np.array_equal(labels.numpy(), G[f'tta_{task}_labels'])
    assert 2 not in labels.tolist()              # the class seen once is dropped (the reference's `<= 1` guard)
    assert np.allclose(boxes.numpy(), G[f'tta_{task}_boxes'], rtol=0, atol=1e-5)
    assert np.allclose(vecs.numpy(), G[f'tta_{task}_vectors'], rtol=0, atol=1e-5)
    for j in (0, 1):                             # per class: sorted by score
        s = boxes[labels == j, 4]
        assert bool((s[:-1] >= s[1:]).all())


def test_vote_cluster_semantics():
    """Hand case: three overlapping boxes (IoU >= 0.66 with the best) + one far away."""
    b = torch.tensor([[10., 10, 50, 50], [11, 10, 51, 50], [10, 12, 50, 52], [100, 100, 140, 140]])
    s = torch.tensor([0.9, 0.6, 0.3, 0.5])
    v = torch.arange(8.).repeat(4, 1) + torch.arange(4.)[:, None]
    ob, ov, os_ = tta.instances_vote(b, v, s)
    # merged box = score-weighted mean, score = best; leftovers 0.6*(1-iou), 0.3*(1-iou) fall below 0.05; far box kept
    w = s[:3] / s[:3].sum()
    assert ob.shape[0] == 2 and torch.allclose(os_, torch.tensor([0.9, 0.5]))
    assert torch.allclose(ob[0], (b[:3] * w[:, None]).sum(0), atol=1e-5) and torch.equal(ob[1], b[3])
    assert torch.allclose(ov[0], (v[:3] * w[:, None]).sum(0), atol=1e-5)
    # a single detection of a class is returned as nothing at all
    e = tta.instances_vote(b[:1], v[:1], s[:1])
    assert e[0].shape == (0, 4) and e[1].shape == (0, 8) and e[2].shape == (0,)


def test_top_k_cut():
    rng = np.random.RandomState(0)
    n = 30
    xy = rng.rand(n, 2) * 1000
    b = torch.from_numpy(np.concatenate([xy, xy + 20, rng.rand(n, 1)], 1).astype(np.float32))      # disjoint boxes
    dets = [(b, torch.zeros(n, 8), torch.zeros(n, dtype=torch.long))]
    meta = [dict(img_shape=(1100, 1100, 3), scale_factor=np.ones(4, np.float32), flip=False)]
    ob, _, _ = tta.vote_merge(dets, meta, 'bbox', 1, 4, [[0, 1e5]], max_per_img=10)
    assert ob.shape[0] == 10 and float(ob[:, 4].min()) >= float(np.sort(b[:, 4].numpy())[-10]) - 1e-7


def test_repeated_augmentation_with_degenerate_boxes():
    """The bookkeeping behind tests/test_gpu_decode.py::test_aug_test_vote_of_a_repeated_augmentation_is_simple_test on
    the host: duplicates of proper boxes merge into themselves; inverted boxes of negative area are filtered; boxes
    with w <= 0 AND h <= 0 (positive 'area') or zero area never overlap anything and stay as two singletons."""
    rng = np.random.RandomState(1)
    n = 40
    xy = rng.rand(n, 2) * 400
    wh = rng.rand(n, 2) * 60 + 1
    wh[:8, 1] *= -1                          # inverted in y: negative area -> filtered
    wh[8:12] *= -1                           # inverted in both: 'area' > 0, IoU with anything 0
    b = np.concatenate([xy, xy + wh, rng.uniform(0.1, 0.9, (n, 1))], 1).astype(np.float32)
    v = rng.rand(n, 8).astype(np.float32)
    l = rng.randint(0, 2, n)
    det = (torch.from_numpy(b), torch.from_numpy(v), torch.from_numpy(l))
    meta = dict(img_shape=(500, 500, 3), scale_factor=np.ones(4, np.float32), flip=False)
    ob, ov, ol = tta.vote_merge([det, det], [meta, meta], 'bbox', 2, 4, [[0, 1e5]])
    ob, ov = ob.numpy(), ov.numpy()
    assert len(ob) == (n - 12) + 2 * 4
    for k in range(n):
        d = np.abs(ob[:, :4] - b[k, :4]).max(1)
        hits = int(((d < 1e-4) & (np.abs(ob[:, 4] - b[k, 4]) < 1e-6)).sum())
        assert hits == (0 if k < 8 else (2 if k < 12 else 1)), (k, hits)
    for k in range(len(ob)):                 # nothing else appears
        assert (np.abs(b[:, :4] - ob[k, :4]).max(1) < 1e-4).any()
