"""Input side of the path against outputs of the UNMODIFIED reference data path recorded in
tests/golden/datapath.npz (tests/golden/make_golden_data.py; inputs rebuilt from tests/golden/synth_coco.py):
annotation parsers, the three train pipelines (single- and multi-scale, seeded flips), contour unification, collate
and the group samplers.  Integer / index outputs: bit-exact.  Coordinates: 1e-4 px.  Normalised pixels: bit-exact."""
import copy
import os
import sys
import types

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import synth_coco as S  # noqa: E402

from lsnet_b200 import datasets as D  # noqa: E402
from lsnet_b200.datasets import contour, loader, transforms  # noqa: E402
from lsnet_b200.registry import DATASETS, PIPELINES  # noqa: E402

G = np.load(os.path.join(HERE, 'golden', 'datapath.npz'))
TASKS = ('bbox', 'segm', 'pose_bbox')


def _dataset(task, ms=False, pipe=None):
    pose = task == 'pose_bbox'
    cls = DATASETS.get('CocoPoseDataset' if pose else 'CocoDataset')
    return cls(ann_file=S.coco_dict(pose), pipeline=pipe if pipe is not None else S.pipeline(task, ms))


def test_registry_names():
    for n in ('LoadImageFromFile', 'LoadAnnotations', 'Resize', 'RandomFlip', 'Normalize', 'Pad', 'DefaultFormatBundle',
              'Collect'):
        assert n in PIPELINES, n
    assert 'CocoDataset' in DATASETS and 'CocoPoseDataset' in DATASETS


@pytest.mark.parametrize('pose', [False, True])
def test_parse_ann_info_matches_reference(pose):
    ds = _dataset('pose_bbox' if pose else 'bbox')
    # the un-annotated image and the one below min_size are filtered (coco.py:95-104); json order is kept
    assert [d['id'] for d in ds.data_infos] == [10 + i for i in range(len(S.SIZES))]
    assert ds.flag.tolist() == [1 if w / h > 1 else 0 for h, w in S.SIZES]
    for i in range(len(ds)):
        ann = ds.get_ann_info(i)
        tag = f'parse_{"pose" if pose else "det"}_{i}'
        assert np.array_equal(ann['bboxes'], G[tag + '_bboxes']) and ann['bboxes'].dtype == np.float32
        assert np.array_equal(ann['labels'], G[tag + '_labels']) and ann['labels'].dtype == np.int64
        assert np.array_equal(ann['bboxes_ignore'], G[tag + '_ignore'])
        key = 'keypoints' if pose else 'extremes'
        assert np.array_equal(ann[key], G[tag + '_' + key])
        assert len(ann['masks']) == int(G[tag + '_nmask'])


def _run(ds, i, ms, img=None):
    info = ds.data_infos[i]
    im = S.image(i) if img is None else img
    res = dict(img_info=info, ann_info=ds.get_ann_info(i), img=im, img_shape=im.shape, ori_shape=im.shape,
               img_fields=['img'], filename=info['filename'], ori_filename=info['filename'])
    ds.pre_pipeline(res)
    np.random.seed(1000 + 10 * i + ms)
    return ds.pipeline(res)


@pytest.mark.parametrize('task', TASKS)
@pytest.mark.parametrize('ms', [0, 1])
def test_pipeline_matches_reference(task, ms):
    ds = _dataset(task, bool(ms))
    flips = 0
    for i in range(len(S.SIZES)):
        out = _run(ds, i, ms)
        tag = f'pipe_{task}_{ms}_{i}'
        meta = out['img_metas']
        got = list(meta['img_shape']) + list(meta['pad_shape']) + [int(meta['flip'])]
        assert got == G[tag + '_meta'].tolist(), (tag, got)        # same scale draw, same flip draw, same padding
        flips += int(meta['flip'])
        assert np.array_equal(np.asarray(meta['scale_factor'], np.float32), G[tag + '_scale_factor'])
        im = out['img'].numpy()
        assert im.dtype == np.float32 and im.shape[0] == 3
        if task == 'bbox':
            ref = G[tag + '_img']
            assert im.shape == ref.shape
            assert np.array_equal(im, ref), float(np.abs(im - ref).max())       # same arithmetic as cv2: bit-exact
        s = np.array([im.astype(np.float64).sum(), np.abs(im.astype(np.float64)).sum()])
        assert np.allclose(s, G[tag + '_imgsum'], rtol=1e-7, atol=1e-3)
        assert np.allclose(out['gt_bboxes'].numpy(), G[tag + '_gt_bboxes'], atol=1e-4, rtol=0)
        assert out['gt_bboxes'].dtype == torch.float32
        assert np.array_equal(out['gt_labels'].numpy(), G[tag + '_gt_labels'])
        if task == 'bbox':
            assert np.allclose(out['gt_extremes'].numpy(), G[tag + '_gt_extremes'], atol=1e-4, rtol=0)
        elif task == 'pose_bbox':
            assert np.allclose(out['gt_keypoints'].numpy(), G[tag + '_gt_keypoints'], atol=1e-4, rtol=0)
        else:
            pm = out['gt_masks']
            assert [pm.height, pm.width] == G[tag + '_mask_hw'].tolist()
            assert [len(c) for c in pm.masks] == G[tag + '_mask_ncomp'].tolist()
            pts = np.concatenate([p for c in pm.masks for p in c])
            assert all(p.shape == (72,) for c in pm.masks for p in c)
            assert np.allclose(pts, G[tag + '_mask_pts'], atol=1e-4, rtol=0), float(np.abs(pts - G[tag + '_mask_pts']).max())
    assert 0 < flips < len(S.SIZES)              # the seeds exercise both branches


def test_head_polygon_step_matches_reference():
    """``LSHead.process_polygons`` (largest component, extent centre appended, extent boxes) on the segm pipeline's
    output — multi-component instances included — against the reference head's own result."""
    from lsnet_b200.modules.head import LSHead
    ds = _dataset('segm', True)
    masks = [_run(ds, i, 1)['gt_masks'] for i in range(len(S.SIZES))]
    assert any(len(c) > 1 for m in masks for c in m.masks)
    polys, boxes = LSHead.process_polygons(None, masks)
    for i, (p, b) in enumerate(zip(polys, boxes)):
        assert p.dtype == torch.float32 and p.shape[1] == 74
        assert np.allclose(p.numpy(), G[f'headpoly_{i}_table'], atol=1e-4, rtol=0)
        assert np.allclose(b.numpy(), G[f'headpoly_{i}_boxes'], atol=1e-4, rtol=0)


def test_uniformsample_vectorised_equals_edge_loop():
    """The array formulation against the reference's per-edge loop (loading.py:311-375) written out for the test:
    random rings on both sides of the 360-point target, including repeated vertices (zero-length edges)."""
    def loop(p, n):
        pn = len(p)
        nxt = p[(np.arange(pn) + 1) % pn]
        el = np.sqrt(((nxt - p) ** 2).sum(1))
        order = np.argsort(el)
        if pn > n:
            return p[np.sort(order[pn - n:])]
        en = contour._edge_budget(el, order, n)
        out = []
        for i in range(pn):
            w = np.arange(en[i], dtype=np.float32).reshape(-1, 1) / en[i]
            out.append(p[i:i + 1] * (1 - w) + nxt[i:i + 1] * w)
        return np.concatenate(out)
    rng = np.random.RandomState(3)
    for n in (3, 4, 7, 50, 359, 360, 361, 500):
        p = rng.rand(n, 2) * 100
        if n >= 7:
            p[3] = p[2]
        a, b = contour.uniformsample(p, 360), loop(p, 360)
        assert a.shape == (360, 2) and np.allclose(a, b, atol=1e-9)


def test_polygon_masks_flip_keeps_start_and_orientation():
    ring = np.array([[5., 1.], [9., 5.], [5., 9.], [1., 5.]])      # clockwise on screen (y down), starts at the top
    pm = contour.PolygonMasks([[ring.reshape(-1)]], 10, 10)
    f = pm.flip('horizontal', keep_cw=True).masks[0][0].reshape(-1, 2)
    assert np.array_equal(f[0], [5., 1.])                          # same start point (mirrored)
    assert np.sign(contour.signed_area(f)) == np.sign(contour.signed_area(ring))
    g = pm.flip('horizontal', keep_cw=False).masks[0][0].reshape(-1, 2)
    assert np.sign(contour.signed_area(g)) == -np.sign(contour.signed_area(ring))


def test_collate_matches_reference():
    ds = _dataset('bbox', True)
    for name, ids in (('a', [0, 3]), ('b', [1, 4, 2])):
        b = loader.collate([_run(ds, i, 1) for i in ids])
        assert list(b['img'].shape) == G[f'collate_{name}_shape'].tolist()
        assert abs(b['img'].double().sum().item() - float(G[f'collate_{name}_sum'])) < 1e-3
        for k, i in enumerate(ids):                      # zero padding bottom / right, content top-left
            ref = G[f'pipe_bbox_1_{i}_img']
            assert np.abs(b['img'][k, :, :ref.shape[1], :ref.shape[2]].numpy() - ref).max() < 1e-5
            assert float(b['img'][k, :, ref.shape[1]:].abs().sum()) == 0 and float(b['img'][k, :, :, ref.shape[2]:].abs().sum()) == 0
        assert len(b['gt_bboxes']) == len(ids) and len(b['img_metas']) == len(ids)


def test_collate_canvas_buckets_keep_valid_extents():
    ds = _dataset('bbox', True)
    dev = _dataset('bbox', True, pipe=loader.device_prep_pipeline(S.pipeline('bbox', True)))
    for d in (ds, dev):
        plain = loader.collate([_run(d, i, 1) for i in (1, 4, 2)])
        b = loader.collate([_run(d, i, 1) for i in (1, 4, 2)], canvas_multiple=128)
        hw = (b['img'].shape[1:3] if b['img'].dtype == torch.uint8 else b['img'].shape[2:])
        assert hw[0] % 128 == 0 and hw[1] % 128 == 0
        assert [m['pad_shape'] for m in b['img_metas']] == [m['pad_shape'] for m in plain['img_metas']]
        ph, pw = (plain['img'].shape[1:3] if plain['img'].dtype == torch.uint8 else plain['img'].shape[2:])
        crop = b['img'][:, :ph, :pw] if b['img'].dtype == torch.uint8 else b['img'][:, :, :ph, :pw]
        assert torch.equal(crop, plain['img'])
        assert int((b['img'] != 0).sum()) == int((plain['img'] != 0).sum())          # the extra canvas is zero


def test_group_samplers_match_reference():
    flag = np.array([1, 0, 1, 1, 0, 1, 1, 0, 0, 1, 1, 1, 0], np.uint8)
    ds = types.SimpleNamespace(flag=flag)
    np.random.seed(5)
    got = list(D.GroupSampler(ds, samples_per_gpu=2))
    assert got == G['group_sampler'].tolist()
    for a, b in zip(got[0::2], got[1::2]):
        assert flag[a] == flag[b]                        # a batch never mixes aspect-ratio groups
    for epoch in (0, 4):
        seen = []
        for rank in range(3):
            s = D.DistributedGroupSampler(ds, samples_per_gpu=2, num_replicas=3, rank=rank)
            s.set_epoch(epoch)
            idx = list(s)
            assert idx == G[f'dist_sampler_e{epoch}_r{rank}'].tolist()
            seen += idx
        assert set(seen) == set(range(len(flag)))        # the ranks cover the dataset


def test_device_prep_pipeline_same_ground_truth_and_pixels():
    """The GPU-side normalisation variant: same ground truth / metas; the uint8 batch + extents reproduce the float
    batch under the kernel's arithmetic (restated in numpy here; the kernel itself is checked in test_gpu_datapath)."""
    ref_ds = _dataset('bbox', True)
    dev_ds = _dataset('bbox', True, pipe=loader.device_prep_pipeline(S.pipeline('bbox', True)))
    ids = [1, 4, 2]
    a = loader.collate([_run(ref_ds, i, 1) for i in ids])
    b = loader.collate([_run(dev_ds, i, 1) for i in ids])
    assert b['img'].dtype == torch.uint8 and list(b['img'].shape) == [3, a['img'].shape[2], a['img'].shape[3], 3]
    for k in range(len(ids)):
        assert torch.equal(a['gt_bboxes'][k], b['gt_bboxes'][k]) and torch.equal(a['gt_extremes'][k], b['gt_extremes'][k])
        for key in ('img_shape', 'pad_shape', 'flip'):
            assert tuple(np.atleast_1d(a['img_metas'][k][key])) == tuple(np.atleast_1d(b['img_metas'][k][key]))
    mean, stdinv = D.DevicePrep.constants(b['img_norm_cfg'])
    u8 = b['img'].numpy()
    x = u8[..., ::-1].astype(np.float64)
    out = ((x - np.array(mean, np.float64)) * np.array(stdinv, np.float64)).astype(np.float32)
    for k, (h, w) in enumerate(b['img_hw'].tolist()):
        out[k, h:] = 0
        out[k, :, w:] = 0
    want = a['img'].numpy().transpose(0, 2, 3, 1)
    assert np.array_equal(out, want)


def test_device_prep_pipeline_rejects_unknown_order():
    with pytest.raises(ValueError):
        loader.device_prep_pipeline([dict(type='Normalize', **S.NORM), dict(type='RandomFlip', flip_ratio=0.5)])
    with pytest.raises(ValueError):
        loader.device_prep_pipeline([dict(type='Pad', size_divisor=32), dict(type='DefaultFormatBundle')])


def test_dataloader_end_to_end(tmp_path):
    """Files on disk -> DataLoader batches -> what the head's ground-truth packing takes."""
    import cv2
    pipe = [dict(type='LoadImageFromFile')] + S.pipeline('segm', True)
    for i in range(len(S.SIZES)):
        cv2.imwrite(str(tmp_path / f'img_{i}.png'), S.image(i))
    cls = DATASETS.get('CocoDataset')
    ds = cls(ann_file=S.coco_dict(False), pipeline=pipe, img_prefix=str(tmp_path))
    dl = D.build_dataloader(ds, samples_per_gpu=2, workers_per_gpu=2, dist=False, seed=3, timeout=180)   # worker processes
    np.random.seed(0)
    from lsnet_b200.data import MODEL_CFG
    from lsnet_b200.registry import build_head
    cfg = MODEL_CFG['segm_r50']
    head = build_head(dict(cfg['model']['bbox_head'], train_cfg=cfg['train_cfg'], test_cfg=None))
    n = 0
    for batch in dl:
        n += 1
        B, _, H, W = batch['img'].shape
        assert B == 2 and H % 32 == 0 and W % 32 == 0
        polys, boxes = head.process_polygons(batch['gt_masks'])
        for P, Bx, gb in zip(polys, boxes, batch['gt_bboxes']):
            assert P.shape == (gb.shape[0], 74) and Bx.shape == (gb.shape[0], 4)
            assert torch.all(Bx[:, 2] > Bx[:, 0]) and torch.all(Bx[:, 3] > Bx[:, 1])
        # ... and the packed ground truth the captured step reads (static capacity, per-level valid extents)
        sizes = [(-(-H // s), -(-W // s)) for s in head.point_strides]
        gt = head.pack_gt(batch['gt_bboxes'], batch['gt_labels'], batch['img_metas'], sizes, 'cpu', capacity=16,
                          gt_masks=batch['gt_masks'])
        assert gt.tables['segm'].shape == (2, 16, 74) and gt.count.tolist() == [len(b) for b in batch['gt_bboxes']]
    assert n == 4            # 3 landscape + 3 portrait/square images, each group padded to whole batches


REF_CFGS = '/root/reference/code/configs/lsnet'


@pytest.mark.skipif(not os.path.isdir(REF_CFGS), reason='reference tree not present')
def test_reference_configs_resolve_their_train_pipelines():
    """``cfg.data.train`` of every LSNet config (the two CPV variants use RepPointsV2 loaders: out of scope)
    names a registered dataset and a pipeline that builds unchanged — and its device-prep rewrite too."""
    import glob
    from lsnet_b200 import Config
    n = 0
    for f in sorted(glob.glob(os.path.join(REF_CFGS, '*.py'))):
        cfg = Config.fromfile(f)
        if cfg.model.bbox_head.type != 'LSHead':
            continue
        tr = cfg.data.train
        assert tr.type in DATASETS, (f, tr.type)
        pipe = D.Compose(tr.pipeline)
        assert [type(t).__name__ for t in pipe.transforms] == [t['type'] for t in tr.pipeline]
        dev = D.Compose(D.device_prep_pipeline(tr.pipeline))
        assert type(dev.transforms[-2]).__name__ == 'DeviceFormatBundle'
        if 'mstrain' in f:
            rs = [t for t in pipe.transforms if isinstance(t, D.Resize)][0]
            assert len(rs.img_scale) == 2 and rs.multiscale_mode == 'range'
        n += 1
    assert n == 15


@pytest.mark.parametrize('multi', [0, 1])
def test_eval_pipeline_matches_reference(multi):
    """MultiScaleFlipAug test pipelines (single scale; two scales x flip): per-augmentation images bit-exact, metas
    equal, and ``collate`` turns a test sample into the lists ``LSDetector.forward_test`` takes."""
    ds = DATASETS.get('CocoDataset')(ann_file=S.coco_dict(False), pipeline=S.eval_pipeline(bool(multi)), test_mode=True)
    assert 'MultiScaleFlipAug' in PIPELINES and 'ImageToTensor' in PIPELINES
    for i in (0, 1):
        info = ds.data_infos[i]
        assert info['id'] == 10 + i                       # test mode keeps every image, in json order
        im = S.image(i)
        res = dict(img_info=info, img=im, img_shape=im.shape, ori_shape=im.shape, img_fields=['img'],
                   filename=info['filename'], ori_filename=info['filename'])
        ds.pre_pipeline(res)
        out = ds.pipeline(res)
        tag = f'test_{multi}_{i}'
        assert len(out['img']) == int(G[tag + '_naug']) == (4 if multi else 1)
        for a, (img, meta) in enumerate(zip(out['img'], out['img_metas'])):
            got = list(meta['img_shape']) + list(meta['pad_shape']) + [int(meta['flip'])]
            assert got == G[f'{tag}_{a}_meta'].tolist()
            assert np.array_equal(np.asarray(meta['scale_factor'], np.float32), G[f'{tag}_{a}_scale_factor'])
            ref = G[f'{tag}_{a}_img']
            if ref.size:
                assert np.array_equal(img.numpy(), ref)
            s = np.array([img.double().sum().item(), img.double().abs().sum().item()])
            assert np.allclose(s, G[f'{tag}_{a}_imgsum'], rtol=1e-9, atol=1e-6)
        b = loader.collate([out])
        assert isinstance(b['img'], list) and len(b['img']) == len(out['img'])
        assert b['img'][0].shape == (1,) + tuple(out['img'][0].shape) and b['img_metas'][0][0] is out['img_metas'][0]


@pytest.mark.skipif(not os.path.isdir(REF_CFGS), reason='reference tree not present')
@pytest.mark.parametrize('cfg,device_prep', [('lsnet_segm_r50_fpn_1x_coco.py', False), ('lsnet_bbox_r50_fpn_1x_coco.py', True)])
def test_train_script_dry_run_on_a_reference_config(tmp_path, capsys, cfg, device_prep):
    """tools/train_coco.py --dry-run: the reference's config file, its dataset / pipeline section unchanged, on image
    files + a COCO json -> the batches the captured step would receive."""
    import importlib.util
    import json
    import cv2
    for i in range(len(S.SIZES)):
        cv2.imwrite(str(tmp_path / f'img_{i}.png'), S.image(i))
    (tmp_path / 'ann.json').write_text(json.dumps(S.coco_dict(False)))
    spec = importlib.util.spec_from_file_location('train_coco', os.path.join(os.path.dirname(HERE), 'tools', 'train_coco.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    argv = [os.path.join(REF_CFGS, cfg), '--ann-file', str(tmp_path / 'ann.json'), '--img-prefix', str(tmp_path),
            '--workers-per-gpu', '0', '--dry-run', '2'] + (['--device-prep'] if device_prep else [])
    np.random.seed(0)
    assert mod.main(argv) == 0
    out = capsys.readouterr().out
    assert 'CocoDataset: 6 images, 4 iterations' in out and 'batch 1: img' in out
    assert ('torch.uint8' in out) == device_prep and ('DeviceFormatBundle' in out) == device_prep
    assert ("'gt_masks'" in out) == ('segm' in cfg) and ("'gt_extremes'" in out) == ('bbox' in cfg)


@pytest.mark.skipif(not os.path.isdir(REF_CFGS), reason='reference tree not present')
def test_test_script_dry_run_on_a_reference_config(tmp_path, capsys):
    """tools/test_coco.py --dry-run: the reference config's own test pipeline (MultiScaleFlipAug), optionally widened to
    several scales + flip, over image files -> the per-image augmentation lists forward_test takes."""
    import importlib.util
    import json
    import cv2
    for i in range(len(S.SIZES)):
        cv2.imwrite(str(tmp_path / f'img_{i}.png'), S.image(i))
    ann = S.coco_dict(False)
    ann['images'] = ann['images'][:len(S.SIZES)]
    (tmp_path / 'ann.json').write_text(json.dumps(ann))
    spec = importlib.util.spec_from_file_location('test_coco', os.path.join(os.path.dirname(HERE), 'tools', 'test_coco.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    base = [os.path.join(REF_CFGS, 'lsnet_bbox_r50_fpn_1x_coco.py'), '--ann-file', str(tmp_path / 'ann.json'),
            '--img-prefix', str(tmp_path), '--workers', '0', '--dry-run', '2']
    assert mod.main(base) == 0
    out = capsys.readouterr().out
    assert 'CocoDataset: 6 images' in out and 'image 1: 1 augmentation(s)' in out and 'flips [False]' in out
    assert mod.main(base + ['--scales', '[(320, 192), (448, 256)]', '--flip']) == 0
    out = capsys.readouterr().out
    assert 'image 0: 4 augmentation(s)' in out and 'flips [False, True, False, True]' in out
