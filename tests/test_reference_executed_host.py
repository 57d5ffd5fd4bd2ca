"""Host logic against the EXECUTED reference (subprocesses: its sys.modules stubs stay out of the other tests).
INTEGRATION.md §1: ``lsnet_b200.registry.install_into_mmdet()`` against the UNMODIFIED reference mmdet /
mmcv (imported from /root/reference through oracle/ref_harness.py, in a subprocess so that its sys.modules stubs stay out of
the other tests).  After the call the reference's own ``build_detector`` / ``build_dataset`` / pipeline ``Compose`` build the
B200 classes from the reference's config file."""
import os
import re
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SECTIONS = {}          # name -> script; all of them run in ONE subprocess (the reference import costs ~15 s)
_RESULT = {}


def _section(name):
    """Run every registered script once (each in its own namespace, failures isolated) and return (ok, output) of one."""
    if not _RESULT:
        parts = ['import sys, traceback, io, contextlib']
        # 'install' swaps the reference's registries for the B200 classes: it must run LAST, the other sections compare
        # against the genuine reference classes
        order = [n for n in _SECTIONS if n != 'install'] + ['install']
        for n in order:
            src = _SECTIONS[n]
            parts.append(f'''
_buf = io.StringIO()
try:
    with contextlib.redirect_stdout(_buf):
        exec(compile({src!r}, {n!r}, 'exec'), {{'__name__': {n!r}}})
    print('@@SECTION {n} OK')
except BaseException:
    print('@@SECTION {n} FAIL')
    traceback.print_exc(file=sys.stdout)
print(_buf.getvalue())
print('@@END {n}')
''')
        r = subprocess.run([sys.executable, '-c', '\n'.join(parts)], capture_output=True, text=True, timeout=1500)
        out = r.stdout
        for n in _SECTIONS:
            m = re.search(rf'@@SECTION {n} (OK|FAIL)\n(.*?)@@END {n}', out, re.S)
            _RESULT[n] = (bool(m) and m.group(1) == 'OK', (m.group(2) if m else out[-2000:]) + r.stderr[-1500:])
    return _RESULT[name]
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
_SECTIONS['install'] = SCRIPT


@pytest.mark.skipif(not os.path.isdir('/root/reference/code/mmdet'), reason='reference tree not present')
def test_install_into_mmdet_swaps_the_reference_registries():
    ok, out = _section('install')
    assert ok and 'INSTALLED' in out, out


LR_SCRIPT = textwrap.dedent('''
    import sys, types, json
    sys.path.insert(0, %r)
    from oracle import ref_harness as rh
    rh.load()
    import torch
    from mmcv.runner.hooks.lr_updater import StepLrUpdaterHook
    from lsnet_b200.train import LrSchedule
    out = {}
    for name, cfg in (('1x', dict(warmup='linear', warmup_iters=500, warmup_ratio=0.001, step=[8, 11])),
                      ('exp', dict(warmup='exp', warmup_iters=40, warmup_ratio=0.1, step=[2], gamma=0.5)),
                      ('const', dict(warmup='constant', warmup_iters=30, warmup_ratio=1.0 / 3, step=[1, 3]))):
        hook = StepLrUpdaterHook(by_epoch=True, **cfg)
        opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.01, momentum=0.9)
        runner = types.SimpleNamespace(optimizer=opt, epoch=0, iter=0)
        hook.before_run(runner)
        ours = LrSchedule(0.01, dict(policy='step', **cfg), iters_per_epoch=60)
        worst = 0.0
        for epoch in range(13):                      # the runner's call order: epoch hook, then one iteration hook per step
            runner.epoch = epoch
            hook.before_train_epoch(runner)
            for i in range(60):
                hook.before_train_iter(runner)
                ref = opt.param_groups[0]['lr']
                worst = max(worst, abs(ref - ours(runner.iter)) / ref)
                runner.iter += 1
        out[name] = worst
    print('LR', json.dumps(out))
''') % ROOT
_SECTIONS['lr'] = LR_SCRIPT


@pytest.mark.skipif(not os.path.isdir('/root/reference/code/mmdet'), reason='reference tree not present')
def test_lr_schedule_matches_the_executed_mmcv_hook():
    """LrSchedule against mmcv's StepLrUpdaterHook driven through the runner's hook order (mmcv/runner/hooks/
    lr_updater.py:100-172) for 13 epochs x 60 iterations: schedule_1x of the LSNet configs plus exp / constant warm-up."""
    import json
    ok, out = _section('lr')
    assert ok, out
    worst = json.loads([l for l in out.splitlines() if l.startswith('LR ')][0][3:])
    assert all(v < 1e-12 for v in worst.values()), worst


PARSE_SCRIPT = textwrap.dedent('''
    import sys, json
    sys.path.insert(0, %r)
    from oracle import ref_harness as rh
    rh.load()
    import torch
    from mmdet.models.detectors.base import BaseDetector
    from lsnet_b200.modules.detector import parse_losses
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.rand(*s, generator=g)
    losses = {'loss_cls': [r(()), r(()), r(()), r(()), r(())], 'loss_bbox_init': [r(3), r(())], 'loss_bbox_refine': r(4, 2),
              'acc': r(()), 'num_pos': [r(())]}
    ref_loss, ref_log = BaseDetector._parse_losses(None, losses)
    loss, log = parse_losses(losses, sync_log=True)
    assert list(log) == list(ref_log)                       # same keys, same order ('acc' / 'num_pos' logged, not summed)
    worst = max(abs(log[k] - ref_log[k]) / abs(ref_log[k]) for k in log)
    worst = max(worst, abs(float(loss) - float(ref_loss)) / float(ref_loss))
    try:
        parse_losses({'loss_x': 1.0})
    except TypeError:
        worst = max(worst, 0.0)
    else:
        worst = 1.0
    print('PARSE', worst)
''') % ROOT
_SECTIONS['parse'] = PARSE_SCRIPT


@pytest.mark.skipif(not os.path.isdir('/root/reference/code/mmdet'), reason='reference tree not present')
def test_parse_losses_matches_the_executed_reference():
    """parse_losses against BaseDetector._parse_losses (mmdet/models/detectors/base.py:176-209) on a loss dict with
    per-level lists, non-scalar entries and logged-only keys."""
    ok, out = _section('parse')
    assert ok, out
    worst = float([l for l in out.splitlines() if l.startswith('PARSE ')][0].split()[1])
    assert worst < 1e-6, worst


CONFIG_SCRIPT = textwrap.dedent('''
    import sys, glob, os
    sys.path.insert(0, %r)
    from oracle import ref_harness as rh
    ns = rh.load()
    import lsnet_b200

    def plain(v):
        if isinstance(v, dict):
            return {k: plain(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return type(v)(plain(x) for x in v)
        return v
    n = 0
    for f in sorted(glob.glob(ns.root + '/configs/lsnet/*.py')):
        ref = plain(ns.Config.fromfile(f)._cfg_dict.to_dict())
        own = plain(lsnet_b200.Config.fromfile(f).to_dict())
        assert own == ref, (os.path.basename(f), sorted(set(own) ^ set(ref)),
                            [k for k in ref if k in own and own[k] != ref[k]])
        n += 1
    print('CONFIGS', n)
''') % ROOT
_SECTIONS['config'] = CONFIG_SCRIPT


@pytest.mark.skipif(not os.path.isdir('/root/reference/code/mmdet'), reason='reference tree not present')
def test_config_loader_equals_mmcv_config_on_every_lsnet_config():
    """lsnet_b200.Config.fromfile against mmcv's Config.fromfile (mmcv/mmcv/utils/config.py: `_base_` inheritance, dict
    merge, `_delete_`): the resulting dictionaries are EQUAL, key for key (tuples stay tuples), for all 17 files under
    configs/lsnet/."""
    ok, out = _section('config')
    assert ok and 'CONFIGS 17' in out, out


INIT_SCRIPT = textwrap.dedent('''
    import sys
    sys.path.insert(0, %r)
    from oracle import ref_harness as rh
    ns = rh.load()
    import torch, lsnet_b200
    import glob, os
    built = 0
    for f in sorted(glob.glob(ns.root + '/configs/lsnet/*.py')):
        name = os.path.basename(f)
        if 'cpv' in name or 'res2' in name:                  # LSCPVDetector / Res2Net: outside the path (DESIGN 7)
            continue
        n_dcn = 30 if 'dconv' in name else 0                 # X-101-DCN: conv2 of the 4 + 23 + 3 blocks of c3-c5
        built += 1
        cfg = ns.Config.fromfile(f); cfg.model.pretrained = None
        ref = ns.build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
        c2 = lsnet_b200.Config.fromfile(f); c2.model.pretrained = None
        own = lsnet_b200.build_detector(c2.model, train_cfg=c2.train_cfg, test_cfg=c2.test_cfg)
        torch.manual_seed(7); ref.init_weights(); a = torch.rand(1).item()
        torch.manual_seed(7); own.init_weights(); b = torch.rand(1).item()
        rs, os_ = ref.state_dict(), own.state_dict()
        assert list(rs) == list(os_), name                                  # same keys in the same order
        assert all(rs[k].shape == os_[k].shape for k in rs), name
        diff = [k for k in rs if not torch.equal(rs[k], os_[k])]
        # the trunk's DCN weights are drawn in the constructor (ResNet.init_weights only re-draws nn.Conv2d): they depend on
        # the construction-time RNG stream, everything else is set by init_weights
        assert len(diff) == n_dcn and all(k.startswith('backbone.') and k.endswith('.conv2.weight') for k in diff), (name, diff[:5])
        assert a == b, name                                                 # init_weights consumed the same random stream
    assert built == 12
    # with the reference's construction-time random stream reproduced (LSNET_REF_INIT_STREAM=1) and the seed set before
    # the build as well, the DCN trunk is identical too
    import os
    os.environ['LSNET_REF_INIT_STREAM'] = '1'
    f = ns.root + '/configs/lsnet/lsnet_bbox_x101_fpn_dconv_c3-c5_mstrain_2x_coco.py.py'
    cfg = ns.Config.fromfile(f); cfg.model.pretrained = None
    torch.manual_seed(3); ref = ns.build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    c2 = lsnet_b200.Config.fromfile(f); c2.model.pretrained = None
    torch.manual_seed(3); own = lsnet_b200.build_detector(c2.model, train_cfg=c2.train_cfg, test_cfg=c2.test_cfg)
    torch.manual_seed(7); ref.init_weights()
    torch.manual_seed(7); own.init_weights()
    rs, os_ = ref.state_dict(), own.state_dict()
    assert list(rs) == list(os_) and all(torch.equal(rs[k], os_[k]) for k in rs), [k for k in rs if not torch.equal(rs[k], os_[k])][:5]
    del os.environ['LSNET_REF_INIT_STREAM']
    assert type(ref).__module__.startswith('mmdet.') and type(own).__module__.startswith('lsnet_b200.')
    print('INIT OK')
''') % ROOT
_SECTIONS['init'] = INIT_SCRIPT


@pytest.mark.skipif(not os.path.isdir('/root/reference/code/mmdet'), reason='reference tree not present')
def test_init_weights_is_bit_identical_to_the_reference():
    """Same seed before ``init_weights()`` -> the same initial model as the reference, tensor for tensor (state_dict
    keys, order, shapes, values) for ALL 12 configs of configs/lsnet/ on this path (R50 and X-101 trunks; bbox, segm,
    pose_bbox, pose_kbox heads; the 5 LSCPVDetector / Res2Net configs are outside it); with DCN in the trunk the 30
    constructor-drawn ``conv2.weight`` tensors are the only ones that differ -- and not even those with
    ``LSNET_REF_INIT_STREAM=1`` and the seed set before the build too."""
    ok, out = _section('init')
    assert ok and 'INIT OK' in out, out


CKPT_SCRIPT = textwrap.dedent('''
    import sys, os, tempfile
    sys.path.insert(0, %r)
    from oracle import ref_harness as rh
    ns = rh.load()
    import torch, lsnet_b200
    from mmcv.runner import load_checkpoint, save_checkpoint as mmcv_save
    from lsnet_b200.train import Trainer, resume, save_checkpoint
    f = ns.root + '/configs/lsnet/lsnet_bbox_r50_fpn_1x_coco.py'
    cfg = ns.Config.fromfile(f); cfg.model.pretrained = None
    tmp = tempfile.mkdtemp()

    def fake_step(params, opt, seed):            # gradients without a forward pass (the kernels need a GPU)
        g = torch.Generator().manual_seed(seed)
        for p in params:
            if p.requires_grad:
                p.grad = torch.randn(p.shape, generator=g) * 1e-3
        opt.step()

    # reference -> here: a checkpoint written by mmcv's save_checkpoint after one SGD step of the reference model
    torch.manual_seed(0)
    ref = ns.build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg); ref.init_weights()
    assert type(ref).__module__.startswith('mmdet.') and type(ref.bbox_head).__module__.startswith('mmdet.')
    ropt = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)      # mmcv build_optimizer: all params
    fake_step(list(ref.parameters()), ropt, 1)
    mmcv_save(ref, os.path.join(tmp, 'ref.pth'), optimizer=ropt, meta=dict(epoch=5, iter=1234))
    c2 = lsnet_b200.Config.fromfile(f); c2.model.pretrained = None
    tr = Trainer(c2, device='cpu')
    meta = resume(tr, os.path.join(tmp, 'ref.pth'))
    assert meta['epoch'] == 5 and tr.iter == 1234
    rs = ref.state_dict()
    assert all(torch.equal(v, rs[k]) for k, v in tr.core.state_dict().items())
    rp, op = list(ref.parameters()), list(tr.core.parameters())
    n = 0
    for a, b in zip(rp, op):
        if a in ropt.state:
            assert torch.equal(ropt.state[a]['momentum_buffer'], tr.optimizer.state[b]['momentum_buffer']); n += 1
        else:
            assert b not in tr.optimizer.state or not tr.optimizer.state[b]
    assert n == sum(p.requires_grad for p in rp) == len(tr.params)
    # ... and both continue identically
    fake_step(rp, ropt, 2); fake_step(op, tr.optimizer, 2)
    assert all(torch.equal(a, b) for a, b in zip(rp, op))

    # here -> reference: our file through mmcv's load_checkpoint + the reference optimizer's load_state_dict
    save_checkpoint(tr, os.path.join(tmp, 'own.pth'), epoch=6)
    torch.manual_seed(9)
    ref2 = ns.build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    ck = load_checkpoint(ref2, os.path.join(tmp, 'own.pth'), map_location='cpu', strict=True)
    assert ck['meta']['epoch'] == 6 and ck['meta']['iter'] == 1234
    ropt2 = torch.optim.SGD(ref2.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    ropt2.load_state_dict(ck['optimizer'])
    assert all(torch.equal(a, b) for a, b in zip(ref2.parameters(), op))
    fake_step(list(ref2.parameters()), ropt2, 3); fake_step(op, tr.optimizer, 3)
    assert all(torch.equal(a, b) for a, b in zip(ref2.parameters(), op))
    print('CKPT OK')
''') % ROOT
_SECTIONS['ckpt'] = CKPT_SCRIPT


@pytest.mark.skipif(not os.path.isdir('/root/reference/code/mmdet'), reason='reference tree not present')
def test_checkpoints_cross_the_boundary_both_ways():
    """A checkpoint written by mmcv's ``save_checkpoint`` (reference model + its SGD over all parameters) resumes here
    -- weights, momentum buffers, epoch / iter -- and both sides take the same next step; a checkpoint written by
    ``lsnet_b200.train.save_checkpoint`` loads through mmcv's ``load_checkpoint(strict=True)`` and the reference
    optimizer's ``load_state_dict`` and continues identically too."""
    ok, out = _section('ckpt')
    assert ok and 'CKPT OK' in out, out


JSON_SCRIPT = textwrap.dedent('''
    import sys, json, os, tempfile
    sys.path.insert(0, %r)
    sys.path.insert(0, %r)
    from oracle import ref_harness as rh
    rh.load()
    import numpy as np
    from mmdet.datasets.coco_pose import CocoPoseDataset as RefPose
    import synth_coco as S
    import lsnet_b200
    rng = np.random.RandomState(0)

    def fake_results(n_img, n_cls, width):
        out = []
        for _ in range(n_img):
            ns = [int(rng.randint(0, 4)) for _ in range(n_cls)]
            xy = [rng.rand(n, 2).astype(np.float32) * 50 for n in ns]
            boxes = [np.concatenate([p, p + 1 + rng.rand(len(p), 2).astype(np.float32) * 30, rng.rand(len(p), 1).astype(np.float32)], 1)
                     for p in xy]
            out.append([boxes, [rng.rand(n, width).astype(np.float32) * 80 for n in ns]])
        return out
    own = lsnet_b200.DATASETS.get('CocoPoseDataset')(ann_file=S.coco_dict(True), pipeline=[], test_mode=True)
    res = fake_results(len(own), 1, 34)

    class Stub:                                  # what the reference methods read from the dataset object
        img_ids, cat_ids = own.img_ids, own.cat_ids
        xyxy2xywh = RefPose.xyxy2xywh
        def __len__(self):
            return len(own)
    assert own._det2json(res) == RefPose._det2json(Stub(), res)
    assert own._kps2json(res) == RefPose._kps2json(Stub(), res)
    tmp = tempfile.mkdtemp()
    files = own.results2json(res, os.path.join(tmp, 'r'))
    assert sorted(files) == ['bbox', 'keypoints', 'proposal'] and json.load(open(files['keypoints'])) == own._kps2json(res)
    det = lsnet_b200.DATASETS.get('CocoDataset')(ann_file=S.coco_dict(False), pipeline=[], test_mode=True)
    r8, r72 = fake_results(len(det), len(det.cat_ids), 8), fake_results(len(det), len(det.cat_ids), 72)
    class Stub2(Stub):
        img_ids, cat_ids = det.img_ids, det.cat_ids
        def __len__(self):
            return len(det)
    assert det._det2json(r8) == RefPose._det2json(Stub2(), r8)
    assert sorted(det.results2json(r8, os.path.join(tmp, 'a'))) == ['bbox', 'proposal']
    f72 = det.results2json(r72, os.path.join(tmp, 'b'))
    segm = json.load(open(f72['segm']))
    assert len(segm) == sum(len(b) for r in r72 for b in r[0]) and all(len(s['segmentation'][0]) == 72 for s in segm)
    print('JSON OK')
''') % (ROOT, os.path.join(ROOT, 'tests', 'golden'))
_SECTIONS['json'] = JSON_SCRIPT


@pytest.mark.skipif(not os.path.isdir('/root/reference/code/mmdet'), reason='reference tree not present')
def test_results_to_coco_json_matches_the_reference():
    """Result lists -> COCO json records: ``_det2json`` / ``_kps2json`` equal to the reference's (coco_pose.py:209-247) on
    random LSNet-format results; the contour variant writes one polygon per instance."""
    ok, out = _section('json')
    assert ok and 'JSON OK' in out, out


MIXED_SCRIPT = textwrap.dedent('''
    import sys, copy, types
    sys.path.insert(0, %r)
    sys.path.insert(0, %r)
    from oracle import ref_harness as rh
    rh.load()
    import numpy as np, torch
    import mmdet.datasets.pipelines.loading as loading
    from mmdet.datasets.builder import PIPELINES as REF
    from mmcv.utils import build_from_cfg
    import make_golden_data as M
    loading.Polygon = M._Polygon
    import synth_coco as S
    import lsnet_b200
    OWN = lsnet_b200.PIPELINES
    assert REF.get('Resize').__module__.startswith('mmdet.')             # the registries have not been swapped yet
    for task in ('bbox', 'segm', 'pose_bbox'):
        ds = lsnet_b200.DATASETS.get('CocoPoseDataset' if task == 'pose_bbox' else 'CocoDataset')(
            ann_file=S.coco_dict(task == 'pose_bbox'), pipeline=[])
        cfgs = S.pipeline(task, True)
        outs = []
        for pick in (lambda i: REF, lambda i: (OWN if i %% 2 else REF), lambda i: (REF if i %% 2 else OWN)):
            stages = [build_from_cfg(dict(c), pick(i)) if pick(i) is REF else lsnet_b200.registry.build_from_cfg(dict(c), pick(i))
                      for i, c in enumerate(cfgs)]
            i = 2
            img = S.image(i)
            r = dict(img_info=ds.data_infos[i], ann_info=copy.deepcopy(ds.get_ann_info(i)), img=img, img_shape=img.shape,
                     ori_shape=img.shape, img_fields=['img'], filename='x', ori_filename='x')
            ds.pre_pipeline(r)
            np.random.seed(5)
            for s_ in stages:
                r = s_(r)
            outs.append(r)
        unwrap = lambda v: v.data if hasattr(v, 'data') and not torch.is_tensor(v) else v
        ref = outs[0]
        for o in outs[1:]:
            assert set(o) == set(ref)
            assert torch.equal(unwrap(o['img']), unwrap(ref['img']))
            for k in ('gt_bboxes', 'gt_labels', 'gt_extremes', 'gt_keypoints'):
                if k in ref:
                    assert torch.equal(unwrap(o[k]), unwrap(ref[k])), (task, k)
            if 'gt_masks' in ref:
                a, b = unwrap(o['gt_masks']), unwrap(ref['gt_masks'])
                assert all(np.array_equal(p, q) for x, y in zip(a.masks, b.masks) for p, q in zip(x, y))
            ma, mb = unwrap(o['img_metas']), unwrap(ref['img_metas'])
            assert ma['img_shape'] == mb['img_shape'] and ma['pad_shape'] == mb['pad_shape'] and ma['flip'] == mb['flip']
    print('MIXED OK')
''') % (ROOT, os.path.join(ROOT, 'tests', 'golden'))
_SECTIONS['mixed'] = MIXED_SCRIPT


@pytest.mark.skipif(not os.path.isdir('/root/reference/code/mmdet'), reason='reference tree not present')
def test_pipeline_stages_interleave_with_the_reference_stages():
    """Every stage reads and writes the reference's ``results`` keys: pipelines that ALTERNATE reference stages and B200
    stages (both phases) give exactly the all-reference output, for the three tasks."""
    ok, out = _section('mixed')
    assert ok and 'MIXED OK' in out, out
