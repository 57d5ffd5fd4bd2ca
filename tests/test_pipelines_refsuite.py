"""The reference's OWN known-answer tests for the input side, re-run against lsnet_b200.datasets
(tests/test_pipelines/test_transform.py: test_resize :14-79, test_flip :82-116, test_pad :217-254, test_normalize
:257-283, test_multi_scale_flip_aug :456-540; tests/test_pipelines/test_formatting.py :8-23; tests/test_masks.py polygon
cases :302-620 — the coordinate-level facts; their bitmap renderings need pycocotools).  The reference reads
tests/data/color.jpg, a 288x512 photo that is not part of the tree: a seeded 288x512 image stands in, every expected
number below (750x1333, 640x1138, 230x409, 345x614 …) is the reference's."""
import copy

import numpy as np
import pytest
import torch

from lsnet_b200.datasets import PolygonMasks
from lsnet_b200.registry import PIPELINES, build_from_cfg


def _img():
    return np.random.RandomState(7).randint(0, 256, (288, 512, 3), dtype=np.uint8)


def _results(extra=True):
    img = _img()
    r = dict(img=img, img_shape=img.shape, ori_shape=img.shape, pad_shape=img.shape, img_fields=['img'])
    if extra:
        r['img2'] = copy.deepcopy(img)
        r['img_fields'] = ['img', 'img2']
    return r


def test_resize():
    with pytest.raises(AssertionError):        # img_scale as a flat list
        build_from_cfg(dict(type='Resize', img_scale=[1333, 800], keep_ratio=True), PIPELINES)
    with pytest.raises(AssertionError):        # several scales together with ratio_range
        build_from_cfg(dict(type='Resize', img_scale=[(1333, 800), (1333, 600)], ratio_range=(0.9, 1.1), keep_ratio=True),
                       PIPELINES)
    with pytest.raises(AssertionError):        # unknown multiscale_mode
        build_from_cfg(dict(type='Resize', img_scale=[(1333, 800), (1333, 600)], keep_ratio=True, multiscale_mode='2333'),
                       PIPELINES)
    resize = build_from_cfg(dict(type='Resize', img_scale=(1333, 800), keep_ratio=True), PIPELINES)
    with pytest.raises(AssertionError):        # scale and scale_factor both preset
        r = _results(False)
        r['scale'], r['scale_factor'] = (1333, 800), 1.0
        resize(r)
    r = resize(_results())
    assert np.equal(r['img'], r['img2']).all() and r['img_shape'] == (750, 1333, 3)
    r.pop('scale'), r.pop('scale_factor')
    r = build_from_cfg(dict(type='Resize', img_scale=(1280, 800), multiscale_mode='value', keep_ratio=False), PIPELINES)(r)
    assert np.equal(r['img'], r['img2']).all() and r['img_shape'] == (800, 1280, 3)


def test_flip():
    with pytest.raises(AssertionError):
        build_from_cfg(dict(type='RandomFlip', flip_ratio=1.5), PIPELINES)
    with pytest.raises(AssertionError):
        build_from_cfg(dict(type='RandomFlip', flip_ratio=1, direction='horizonta'), PIPELINES)
    cfg = dict(type='RandomFlip', flip_ratio=1)
    r = _results()
    r['scale_factor'] = 1.0
    original = copy.deepcopy(r['img'])
    r = build_from_cfg(cfg, PIPELINES)(r)
    assert np.equal(r['img'], r['img2']).all() and not np.equal(r['img'], original).all()
    r = build_from_cfg(cfg, PIPELINES)(r)       # 'flip' is now preset True: flips back
    assert np.equal(r['img'], r['img2']).all() and np.equal(original, r['img']).all()


def test_pad():
    with pytest.raises(AssertionError):
        build_from_cfg(dict(type='Pad'), PIPELINES)
    pad = build_from_cfg(dict(type='Pad', size_divisor=32), PIPELINES)
    r = _results()
    r['scale_factor'] = 1.0
    original = copy.deepcopy(r['img'])
    r = pad(r)
    assert np.equal(r['img'], r['img2']).all() and np.equal(r['img'], original).all()      # 288x512 is divisible already
    r.pop('scale_factor')
    r = pad(build_from_cfg(dict(type='Resize', img_scale=(1333, 800), keep_ratio=True), PIPELINES)(r))
    assert np.equal(r['img'], r['img2']).all()
    assert r['img'].shape[0] % 32 == 0 and r['img'].shape[1] % 32 == 0 and r['pad_shape'] == (768, 1344, 3)
    assert not r['img'][750:].any() and not r['img'][:, 1333:].any()


def test_normalize():
    cfg = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
    r = _results()
    original = copy.deepcopy(r['img'])
    r = build_from_cfg(dict(type='Normalize', **cfg), PIPELINES)(r)
    assert np.equal(r['img'], r['img2']).all() and r['img'].dtype == np.float32
    assert np.allclose(r['img'], (original[..., ::-1] - np.array(cfg['mean'])) / np.array(cfg['std']))
    assert r['img_norm_cfg']['to_rgb'] is True


def test_multi_scale_flip_aug():
    inner = [dict(type='Resize')]
    for bad in (dict(scale_factor=1.0, img_scale=[(1333, 800)]), dict(scale_factor=None, img_scale=None),
                dict(img_scale=[1333, 800]), dict(img_scale=[(1333, 800)], flip_direction=1)):
        with pytest.raises(AssertionError):
            build_from_cfg(dict(type='MultiScaleFlipAug', transforms=inner, **bad), PIPELINES)
    r = _results(False)
    t = build_from_cfg(dict(type='MultiScaleFlipAug', img_scale=[(1333, 800), (1333, 640)],
                            transforms=[dict(type='Resize', keep_ratio=True)]), PIPELINES)
    out = t(copy.deepcopy(r))
    assert len(out['img']) == 2
    assert out['img'][0].shape == (750, 1333, 3) and out['img_shape'][0] == (750, 1333, 3)
    assert out['img'][1].shape == (640, 1138, 3) and out['img_shape'][1] == (640, 1138, 3)
    t = build_from_cfg(dict(type='MultiScaleFlipAug', scale_factor=[0.8, 1.0, 1.2],
                            transforms=[dict(type='Resize', keep_ratio=False)]), PIPELINES)
    out = t(copy.deepcopy(r))
    assert [im.shape for im in out['img']] == [(230, 409, 3), (288, 512, 3), (345, 614, 3)]
    assert out['img_shape'] == [(230, 409, 3), (288, 512, 3), (345, 614, 3)]
    # the test pipeline of configs/_base_/datasets/coco_detection.py:16-29
    norm = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
    t = build_from_cfg(dict(type='MultiScaleFlipAug', img_scale=(1333, 800), flip=False, transforms=[
        dict(type='Resize', keep_ratio=True), dict(type='RandomFlip'), dict(type='Normalize', **norm),
        dict(type='Pad', size_divisor=32), dict(type='ImageToTensor', keys=['img']), dict(type='Collect', keys=['img'])]),
        PIPELINES)
    out = t(dict(copy.deepcopy(r), filename='color.jpg', ori_filename='color.jpg'))
    assert len(out['img']) == 1 and len(out['img_metas']) == 1 and isinstance(out['img'][0], torch.Tensor)
    assert tuple(out['img'][0].shape) == (3, 768, 1344) and out['img_metas'][0]['flip'] is False


def test_default_format_bundle_adds_default_meta_keys():
    img = _img()
    r = dict(img=img, img_shape=img.shape, ori_shape=img.shape, img_fields=['img'])
    assert 'pad_shape' not in r and 'scale_factor' not in r and 'img_norm_cfg' not in r
    r = build_from_cfg(dict(type='DefaultFormatBundle'), PIPELINES)(r)
    assert r['pad_shape'] == (288, 512, 3) and r['scale_factor'] == 1.0 and r['img_norm_cfg']['to_rgb'] is False
    assert isinstance(r['img'], torch.Tensor) and tuple(r['img'].shape) == (3, 288, 512)


def _polys(n, rng):
    return [[rng.uniform(0, 28, int(rng.randint(5)) * 2 + 6)] for _ in range(n)]


def test_polygon_masks():
    rng = np.random.RandomState(0)
    # init / len / type checks (test_masks.py:302-327)
    pm = PolygonMasks(_polys(3, rng), 28, 28)
    assert len(pm) == 3 and pm.height == 28 and pm.width == 28 and isinstance(pm.masks[0][0], np.ndarray)
    with pytest.raises(AssertionError):
        PolygonMasks([[[]]], 28, 28)
    # rescale: the new size is the image's rescaled size, empty or not (:330-355)
    e = PolygonMasks([], 28, 28).rescale((56, 72))
    assert len(e) == 0 and (e.height, e.width) == (56, 56)
    one = PolygonMasks([[np.array([1, 1, 3, 1, 4, 3, 2, 4, 1, 3], dtype=float)]], 5, 5)
    r = one.rescale((12, 10))
    assert len(r) == 1 and (r.height, r.width) == (10, 10)
    assert np.array_equal(r.masks[0][0], [2, 2, 6, 2, 8, 6, 4, 8, 2, 6])
    # resize: per-axis factors, several parts, several instances (:358-410)
    e = PolygonMasks([], 28, 28).resize((56, 72))
    assert len(e) == 0 and (e.height, e.width) == (56, 72)
    two = PolygonMasks([[np.array([0., 0., 1., 0., 1., 1.]), np.array([1., 1., 2., 1., 2., 2., 1., 2.])]], 3, 3)
    r = two.resize((6, 12))
    assert np.array_equal(r.masks[0][0], [0, 0, 4, 0, 4, 2]) and np.array_equal(r.masks[0][1], [4, 2, 8, 2, 8, 4, 4, 4])
    assert np.array_equal(two.masks[0][0], [0., 0., 1., 0., 1., 1.])            # the source is not modified
    # flip twice = identity, both directions (:413-446)
    for d in ('horizontal', 'vertical'):
        f = pm.flip(d)
        assert len(f) == 3 and (f.height, f.width) == (28, 28)
        ff = f.flip(d)
        assert all(np.allclose(a[0], b[0]) for a, b in zip(pm.masks, ff.masks))
        assert not np.allclose(pm.masks[0][0], f.masks[0][0])
    # pad only changes the canvas (:478-496)
    p = pm.pad((56, 56))
    assert len(p) == 3 and (p.height, p.width) == (56, 56) and p.masks is pm.masks
    assert len(PolygonMasks([], 28, 28).pad((56, 56))) == 0
    # areas: shoelace; a 4 x 3 triangle (:531-545)
    assert PolygonMasks([], 28, 28).areas.sum() == 0
    assert np.isclose(PolygonMasks([[np.array([1, 1, 5, 1, 3, 4])]], 6, 6).areas, [6.0]).all()
    # indexing and iteration (:589-606)
    assert len(pm[0]) == 1 and len(pm[[0, 1]]) == 2 and len(pm[np.asarray([0, 1])]) == 2
    with pytest.raises(ValueError):
        pm[torch.Tensor([1, 2])]
    for i, m in enumerate(pm):
        assert np.equal(m, pm.masks[i]).all()


@pytest.mark.parametrize('name', ['CocoDataset', 'CocoPoseDataset'])
def test_custom_classes_override_default(name, tmp_path):
    """tests/test_dataset.py:15-86 (on a real annotation dict instead of MagicMocks)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import synth_coco as S
    from lsnet_b200.registry import DATASETS
    cls = DATASETS.get(name)
    original = cls.CLASSES
    ann = S.coco_dict(name == 'CocoPoseDataset')
    d = cls(ann_file=ann, pipeline=[], classes=('bus', 'car'), test_mode=True)
    assert d.CLASSES != original and d.CLASSES == ('bus', 'car') and d.custom_classes
    d = cls(ann_file=ann, pipeline=[], classes=['bus', 'car'], test_mode=True)
    assert d.CLASSES == ['bus', 'car'] and d.custom_classes
    d = cls(ann_file=ann, pipeline=[], classes=['foo'], test_mode=True)
    assert d.CLASSES == ['foo'] and d.custom_classes and len(d) == 0          # no image holds a 'foo'
    d = cls(ann_file=ann, pipeline=[], classes=None, test_mode=True)
    assert d.CLASSES == original and not d.custom_classes and cls.CLASSES == original
    f = tmp_path / 'classes.txt'
    f.write_text('bus\ncar\n')
    d = cls(ann_file=ann, pipeline=[], classes=str(f), test_mode=True)
    assert d.CLASSES == ['bus', 'car'] and d.custom_classes
    if name == 'CocoDataset':
        # the subset: images with at least one 'car' (category 3), labels re-indexed over the custom classes
        d = cls(ann_file=ann, pipeline=[], classes=('car',))
        has_car = sorted({a['image_id'] for a in ann['annotations'] if a['category_id'] == 3})
        assert [i['id'] for i in d.data_infos] == [i for i in has_car if i < 90] and d.cat2label == {3: 0}
        assert all((d.get_ann_info(k)['labels'] == 0).all() for k in range(len(d)))
