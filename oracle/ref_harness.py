"""ORACLE — test infrastructure only (see oracle/__init__.py).

Imports the *unmodified* reference Python (mmdet + vendored mmcv 0.6.2) from
/root/reference/code in THIS container so that the restatement in
``lsnet_oracle.py`` can be validated against it and golden vectors generated
(tests/golden/make_golden.py).  /root/reference does not exist on the GPU box:
nothing that runs there may import this module.

What is stubbed: compiled extensions and optional third-party
modules that the hot path never calls.  What is *replaced*: the CUDA-only
``deform_conv_ext`` / ``sigmoid_focal_loss_ext`` entry points, by the CPU
restatement in ``dcn_ops.py`` plugged in underneath the reference's own
autograd Functions; and PointGenerator's ``device='cuda'`` default.
"""
import importlib.machinery
import os
import sys
import types

REF_ROOT = '/root/reference/code'


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'mmdet'))


class _Lenient(types.ModuleType):
    """Stub module: any missing attribute resolves to an inert placeholder class."""

    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return type(k, (object,), {'__init__': lambda self, *a, **kw: None,
                                   '__call__': lambda self, *a, **kw: None,
                                   '__class_getitem__': classmethod(lambda cls, i: cls)}) \
            if k[:1].isupper() else (lambda *a, **kw: None)


def _stub(name, **attrs):
    m = _Lenient(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _AddictDict(dict):
    """Minimal addict.Dict (attribute access, recursive wrap) for mmcv.Config."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for a in args:
            if isinstance(a, dict):
                for k, v in a.items():
                    self[k] = self._hook(v)
        for k, v in kwargs.items():
            self[k] = self._hook(v)

    @classmethod
    def _hook(cls, v):
        if isinstance(v, dict) and not isinstance(v, cls):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._hook(i) for i in v)
        return v

    def __setattr__(self, k, v):
        self[k] = v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._hook(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __missing__(self, k):
        raise KeyError(k)

    def to_dict(self):
        out = {}
        for k, v in self.items():
            if isinstance(v, _AddictDict):
                out[k] = v.to_dict()
            elif isinstance(v, (list, tuple)):
                out[k] = type(v)(i.to_dict() if isinstance(i, _AddictDict) else i for i in v)
            else:
                out[k] = v
        return out

    def __deepcopy__(self, memo):
        import copy
        return type(self)({k: copy.deepcopy(v, memo) for k, v in self.items()})


_LOADED = {}


def load():
    """Import the reference; returns a namespace with the modules we need."""
    if _LOADED:
        return _LOADED['ns']
    if not available():
        raise RuntimeError('reference tree not present (expected on the GPU box)')
    _stub('addict', Dict=_AddictDict)
    _stub('yapf')
    _stub('yapf.yapflib')
    _stub('yapf.yapflib.yapf_api', FormatCode=lambda s, **kw: (s, False))
    for name in ['mmcv._ext', 'matplotlib', 'matplotlib.pyplot', 'terminaltables', 'pycocotools',
                 'pycocotools.mask', 'pycocotools.coco', 'pycocotools.cocoeval', 'shapely',
                 'shapely.geometry']:
        _stub(name)
    sys.modules['terminaltables'].AsciiTable = object
    sys.modules['pycocotools.coco'].COCO = object
    sys.modules['pycocotools.cocoeval'].COCOeval = object
    sys.modules['pycocotools'].mask = sys.modules['pycocotools.mask']
    sys.modules['shapely.geometry'].Polygon = object
    sys.path[:0] = [os.path.join(REF_ROOT, 'mmcv'), REF_ROOT]
    _stub('mmdet.version', __version__='2.1.0+ref', short_version='2.1.0')
    for ext in ['utils.compiling_info', 'nms.nms_ext', 'roi_align.roi_align_ext', 'roi_pool.roi_pool_ext',
                'dcn.deform_conv_ext', 'dcn.deform_pool_ext', 'sigmoid_focal_loss.sigmoid_focal_loss_ext',
                'masked_conv.masked_conv2d_ext', 'carafe.carafe_ext', 'carafe.carafe_naive_ext',
                'corner_pool.corner_pool_ext', 'chamfer_2d.chamfer_2d']:
        _stub('mmdet.ops.' + ext)
    ci = sys.modules['mmdet.ops.utils.compiling_info']
    ci.get_compiler_version = lambda: 'stub'
    ci.get_compiling_cuda_version = lambda: 'stub'

    import torch
    import mmcv  # noqa: F401
    import mmdet.models  # noqa: F401
    import mmdet.core  # noqa: F401
    from mmcv import Config
    from mmdet.models import build_detector

    from . import dcn_ops

    # --- plug the CPU restatement under the reference's own autograd Functions -----------------
    dc = sys.modules['mmdet.ops.dcn.deform_conv']

    # The reference Functions raise NotImplementedError on CPU tensors (deform_conv.py:46,136,221),
    # so the module-level callables are re-pointed at the oracle ops, which take the same arguments.
    dc.deform_conv = dcn_ops.deform_conv
    dc.modulated_deform_conv = dcn_ops.modulated_deform_conv
    dc.pyramid_deform_conv = dcn_ops.pyramid_deform_conv

    fl = sys.modules['mmdet.ops.sigmoid_focal_loss.sigmoid_focal_loss']
    fl.sigmoid_focal_loss = dcn_ops.sigmoid_focal_loss_elementwise
    sys.modules['mmdet.ops.sigmoid_focal_loss'].sigmoid_focal_loss = dcn_ops.sigmoid_focal_loss_elementwise
    sys.modules['mmdet.ops'].sigmoid_focal_loss = dcn_ops.sigmoid_focal_loss_elementwise
    mfl = sys.modules['mmdet.models.losses.focal_loss']
    mfl._sigmoid_focal_loss = dcn_ops.sigmoid_focal_loss_elementwise

    # --- PointGenerator default device (point_generator.py:17,27) -----------------------------
    from mmdet.core.anchor.point_generator import PointGenerator
    _gp, _vf = PointGenerator.grid_points, PointGenerator.valid_flags
    PointGenerator.grid_points = lambda self, fs, stride=16, device='cpu': _gp(self, fs, stride, device)
    PointGenerator.valid_flags = lambda self, fs, vs, device='cpu': _vf(self, fs, vs, device)

    ns = types.SimpleNamespace(torch=torch, Config=Config, build_detector=build_detector,
                               mmdet=sys.modules['mmdet'], dc=dc, root=REF_ROOT)
    _LOADED['ns'] = ns
    return ns


def build_reference_detector(cfg_name='lsnet_bbox_r50_fpn_1x_coco.py', overrides=None):
    ns = load()
    cfg = ns.Config.fromfile(os.path.join(REF_ROOT, 'configs', 'lsnet', cfg_name))
    cfg.model.pretrained = None
    if overrides:
        overrides(cfg)
    model = ns.build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    return model, cfg
