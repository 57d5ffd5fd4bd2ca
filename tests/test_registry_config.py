"""The drop-in boundary on CPU: registry semantics (mmcv/tests/test_registry.py), config inheritance, every shipped
LSNet R50 config builds with the reference's parameter names, and the product fails loudly without a GPU."""
import glob
import os

import pytest
import torch

import lsnet_b200 as L
from lsnet_b200.registry import Registry, build_from_cfg
from oracle import init as oinit

REF_CFGS = '/root/reference/code/configs/lsnet'


def test_registry_semantics():
    r = Registry('cat')

    @r.register_module()
    class A:
        def __init__(self, x=1):
            self.x = x
    with pytest.raises(KeyError):
        r.register_module(module=A)
    r.register_module(module=A, force=True)
    r.register_module(name='alias', module=A)
    assert r.get('alias') is A and 'A' in r and len(r) == 2
    assert build_from_cfg(dict(type='A', x=3), r).x == 3
    assert build_from_cfg(dict(type=A), r, default_args=dict(x=5)).x == 5
    assert build_from_cfg(dict(type='A', x=3), r, default_args=dict(x=5)).x == 3
    with pytest.raises(KeyError):
        build_from_cfg(dict(type='B'), r)
    with pytest.raises(KeyError):
        build_from_cfg(dict(x=1), r)
    with pytest.raises(TypeError):
        build_from_cfg([], r)
    with pytest.raises(TypeError):
        r.register_module(module=3)


def test_names_registered():
    for reg, names in ((L.DETECTORS, ['LSDetector']), (L.BACKBONES, ['ResNet', 'ResNeXt']), (L.NECKS, ['FPN']),
                       (L.HEADS, ['LSHead']), (L.LOSSES, ['CrossIOULoss', 'FocalLoss']),
                       (L.BBOX_ASSIGNERS, ['CentroidAssigner', 'ATSSAssigner']), (L.CONV_LAYERS, ['DCN', 'DCNv2'])):
        for n in names:
            assert reg.get(n) is not None, n


def test_config_base_and_delete(tmp_path):
    (tmp_path / 'base.py').write_text("optimizer = dict(type='SGD', lr=0.02)\noptimizer_config = dict(grad_clip=None)\n"
                                      "model = dict(a=dict(b=1, c=2))\n")
    (tmp_path / 'child.py').write_text("_base_ = ['./base.py']\noptimizer = dict(lr=0.01)\n"
                                       "optimizer_config = dict(grad_clip=dict(max_norm=35), _delete_=True)\n"
                                       "model = dict(a=dict(b=5))\n")
    c = L.Config.fromfile(str(tmp_path / 'child.py'))
    assert c.optimizer == dict(type='SGD', lr=0.01)
    assert c.optimizer_config == dict(grad_clip=dict(max_norm=35))
    assert c.model.a.b == 5 and c.model.a.c == 2


@pytest.mark.skipif(not os.path.isdir(REF_CFGS), reason='reference tree not present')
def test_every_shipped_lsnet_config_loads_and_r50_builds():
    files = sorted(glob.glob(os.path.join(REF_CFGS, '*.py')))
    assert len(files) == 17
    for f in files:
        c = L.Config.fromfile(f)
        assert c.model.bbox_head.type in ('LSHead', 'LSCPVHead')
        if c.model.backbone.type == 'ResNet' and c.model.type == 'LSDetector':
            c.model.pretrained = None
            m = L.build_detector(c.model, train_cfg=c.train_cfg, test_cfg=c.test_cfg)
            task = c.model.bbox_head.task
            ref_keys = set(oinit.make_state_dict(task, 0).keys())
            assert set(m.state_dict().keys()) == ref_keys, f


def test_parameter_counts_match_survey():
    from lsnet_b200.data import MODEL_CFG
    cfg = MODEL_CFG['bbox_r50']
    m = L.build_detector(cfg['model'], train_cfg=cfg['train_cfg'])
    assert sum(p.numel() for p in m.parameters()) == 38802018
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 38576674
    # conv_offset zero-init (P6) and the cls prior bias
    assert float(m.bbox_head.cls_convs[0].conv.conv_offset.weight.abs().max()) == 0.0
    assert abs(float(m.bbox_head.pts_cls_out.bias[0]) + 4.59512) < 1e-4


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import lsnet_b200.ops as ops
    from lsnet_b200.lib import LsnetError
    x = torch.randn(1, 64, 8, 8)
    w = torch.randn(32, 64, 3, 3)
    with pytest.raises((LsnetError, AssertionError, RuntimeError)):
        ops.conv2d_same(x, w, None, padding=1)


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, 'lsnet_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f
