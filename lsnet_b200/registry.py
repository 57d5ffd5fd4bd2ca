"""The drop-in boundary: string -> class registries with the semantics of the reference's
``mmcv.utils.Registry`` / ``build_from_cfg`` (mmcv/mmcv/utils/registry.py:8-167): duplicate names raise KeyError
unless ``force=True``; a cfg dict needs a ``type`` key that is a registered name or a class; ``default_args`` only
fill missing keys.  The registries carry the reference's names so the shipped configs resolve:
mmdet/models/builder.py:4-10 (BACKBONES, NECKS, HEADS, LOSSES, DETECTORS), mmdet/core/bbox/builder.py:3-5
(BBOX_ASSIGNERS, BBOX_SAMPLERS), mmcv/mmcv/cnn/bricks/registry.py (CONV_LAYERS).

If the real mmcv/mmdet are importable, ``install_into_mmdet()`` registers the B200 modules into THEIR registries with
force=True, which is the one-line switch a reference user makes.
"""
import inspect


class Registry:

    def __init__(self, name):
        self._name = name
        self._modules = {}

    name = property(lambda self: self._name)
    module_dict = property(lambda self: self._modules)

    def __len__(self):
        return len(self._modules)

    def __contains__(self, key):
        return key in self._modules

    def __repr__(self):
        return f'Registry(name={self._name}, items={sorted(self._modules)})'

    def get(self, key):
        return self._modules.get(key)

    def _add(self, cls, name, force):
        if not inspect.isclass(cls):
            raise TypeError(f'module must be a class, but got {type(cls)}')
        name = name or cls.__name__
        if name in self._modules and not force:
            raise KeyError(f'{name} is already registered in {self._name}')
        self._modules[name] = cls

    def register_module(self, name=None, force=False, module=None):
        if not isinstance(force, bool):
            raise TypeError(f'force must be a boolean, but got {type(force)}')
        if inspect.isclass(name):            # old positional-class form: registry.register_module(Cls)
            self._add(name, None, force)
            return name
        if not (name is None or isinstance(name, str)):
            raise TypeError(f'name must be a str, but got {type(name)}')
        if module is not None:
            self._add(module, name, force)
            return module

        def deco(cls):
            self._add(cls, name, force)
            return cls
        return deco


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict):
        raise TypeError(f'cfg must be a dict, but got {type(cfg)}')
    if 'type' not in cfg:
        raise KeyError(f'the cfg dict must contain the key "type", but got {cfg}')
    if not isinstance(registry, Registry):
        raise TypeError(f'registry must be a Registry object, but got {type(registry)}')
    if not (default_args is None or isinstance(default_args, dict)):
        raise TypeError(f'default_args must be a dict or None, but got {type(default_args)}')
    kwargs = dict(cfg)
    kind = kwargs.pop('type')
    if isinstance(kind, str):
        cls = registry.get(kind)
        if cls is None:
            raise KeyError(f'{kind} is not in the {registry.name} registry')
    elif inspect.isclass(kind):
        cls = kind
    else:
        raise TypeError(f'type must be a str or valid type, but got {type(kind)}')
    for k, v in (default_args or {}).items():
        kwargs.setdefault(k, v)
    return cls(**kwargs)


BACKBONES = Registry('backbone')
NECKS = Registry('neck')
HEADS = Registry('head')
LOSSES = Registry('loss')
DETECTORS = Registry('detector')
BBOX_ASSIGNERS = Registry('bbox_assigner')
BBOX_SAMPLERS = Registry('bbox_sampler')
CONV_LAYERS = Registry('conv layer')
DATASETS = Registry('dataset')          # mmdet/datasets/builder.py:10-11
PIPELINES = Registry('pipeline')


def build(cfg, registry, default_args=None):
    """mmdet/models/builder.py:13-32: a list of cfgs builds an nn.Sequential."""
    if isinstance(cfg, (list, tuple)):
        import torch.nn as nn
        return nn.Sequential(*[build_from_cfg(c, registry, default_args) for c in cfg])
    return build_from_cfg(cfg, registry, default_args)


def build_backbone(cfg):
    return build(cfg, BACKBONES)


def build_neck(cfg):
    return build(cfg, NECKS)


def build_head(cfg):
    return build(cfg, HEADS)


def build_loss(cfg):
    return build(cfg, LOSSES)


def build_assigner(cfg, **default_args):
    return build_from_cfg(cfg, BBOX_ASSIGNERS, default_args)


def build_sampler(cfg, **default_args):
    return build_from_cfg(cfg, BBOX_SAMPLERS, default_args)


def build_detector(cfg, train_cfg=None, test_cfg=None):
    """mmdet/models/builder.py:65-67."""
    return build(cfg, DETECTORS, dict(train_cfg=train_cfg, test_cfg=test_cfg))


def build_conv_layer(cfg, *args, **kwargs):
    """mmcv/mmcv/cnn/bricks/conv.py:11-43: cfg None -> plain Conv2d; else CONV_LAYERS[type](*args, **kwargs, **cfg)."""
    import torch.nn as nn
    if cfg is None:
        cfg_ = dict(type='Conv')
    else:
        if not isinstance(cfg, dict):
            raise TypeError('cfg must be a dict')
        if 'type' not in cfg:
            raise KeyError('the cfg dict must contain the key "type"')
        cfg_ = dict(cfg)
    kind = cfg_.pop('type')
    if kind in ('Conv', 'Conv2d'):
        return nn.Conv2d(*args, **kwargs, **cfg_)
    cls = CONV_LAYERS.get(kind)
    if cls is None:
        raise KeyError(f'Unrecognized norm type {kind}')
    return cls(*args, **kwargs, **cfg_)


def install_into_mmdet():
    """Register the B200 modules into an importable mmdet/mmcv (force=True) so existing configs pick them up:
    models, assigners / samplers, conv layers and -- when mmdet.datasets is importable -- datasets and pipelines."""
    from mmcv.cnn import CONV_LAYERS as M_CONV
    from mmdet.core.bbox.builder import BBOX_ASSIGNERS as M_ASSIGN, BBOX_SAMPLERS as M_SAMPLE
    from mmdet.models.builder import BACKBONES as MB, DETECTORS as MD, HEADS as MH, LOSSES as ML, NECKS as MN
    pairs = [(BACKBONES, MB), (NECKS, MN), (HEADS, MH), (LOSSES, ML), (DETECTORS, MD), (BBOX_ASSIGNERS, M_ASSIGN),
             (BBOX_SAMPLERS, M_SAMPLE), (CONV_LAYERS, M_CONV)]
    try:
        from mmdet.datasets.builder import DATASETS as M_DATA, PIPELINES as M_PIPE
        pairs += [(DATASETS, M_DATA), (PIPELINES, M_PIPE)]
    except ImportError:          # mmdet.datasets needs pycocotools; the model side does not
        pass
    for src, dst in pairs:
        for name, cls in src.module_dict.items():
            dst.register_module(name=name, force=True, module=cls)
