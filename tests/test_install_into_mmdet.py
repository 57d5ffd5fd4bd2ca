"""INTEGRATION.md §1, executed: ``lsnet_b200.registry.install_into_mmdet()`` against the UNMODIFIED reference mmdet /
mmcv (imported from /root/reference through oracle/ref_harness.py, in a subprocess so that its sys.modules stubs stay out of
the other tests).  After the call the reference's own ``build_detector`` / ``build_dataset`` / pipeline ``Compose`` build the
B200 classes from the reference's config file."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = textwrap.dedent('''
    import sys
    sys.path.insert(0, %r)
    from oracle import ref_harness as rh
    ns = rh.load()
    import mmdet.datasets                                    # the reference's registries, all of them
    from mmdet.models.builder import DETECTORS, HEADS, BACKBONES
    from mmdet.datasets.builder import PIPELINES, DATASETS
    ref_head, ref_resize = HEADS.get('LSHead'), PIPELINES.get('Resize')
    import lsnet_b200
    lsnet_b200.registry.install_into_mmdet()
    assert HEADS.get('LSHead') is lsnet_b200.HEADS.get('LSHead') is not ref_head
    assert PIPELINES.get('Resize') is lsnet_b200.PIPELINES.get('Resize') is not ref_resize
    assert DATASETS.get('CocoDataset') is lsnet_b200.DATASETS.get('CocoDataset')
    cfg = ns.Config.fromfile(ns.root + '/configs/lsnet/lsnet_bbox_r50_fpn_1x_coco.py')
    cfg.model.pretrained = None
    model = ns.build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)     # mmdet's own builder
    assert type(model).__module__.startswith('lsnet_b200.') and type(model.bbox_head).__module__.startswith('lsnet_b200.')
    assert type(model.backbone).__module__.startswith('lsnet_b200.') and type(model.neck).__module__.startswith('lsnet_b200.')
    assert type(model.bbox_head.cls_convs[0].conv).__module__.startswith('lsnet_b200.')      # CONV_LAYERS 'DCNv2'
    assert sum(p.numel() for p in model.parameters()) == 38802018                           # SURVEY 8c
    from mmdet.datasets.pipelines import Compose
    pipe = Compose(cfg.data.train.pipeline)                                                 # mmdet's own Compose
    assert all(type(t).__module__.startswith('lsnet_b200.') for t in pipe.transforms)
    print('INSTALLED')
''') % ROOT


@pytest.mark.skipif(not os.path.isdir('/root/reference/code/mmdet'), reason='reference tree not present')
def test_install_into_mmdet_swaps_the_reference_registries():
    r = subprocess.run([sys.executable, '-c', SCRIPT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'INSTALLED' in r.stdout, r.stderr[-3000:]
