"""Loader for the reference's Python-file configs (mmcv/mmcv/utils/config.py:90-222): ``_base_`` (str or list)
inheritance with recursive dict merge, ``_delete_=True`` to replace instead of merge, attribute access on nested
dicts.  Enough to load configs/lsnet/*.py unmodified; no yapf / pretty-printing / json-yaml support."""
import os
import runpy

BASE_KEY = '_base_'
DELETE_KEY = '_delete_'


class ConfigDict(dict):
    """dict with attribute access (the subset of addict.Dict the configs rely on)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(f"'ConfigDict' object has no attribute '{name}'")

    def __setattr__(self, name, value):
        self[name] = value

    def __deepcopy__(self, memo):
        import copy
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _wrap(v):
    if isinstance(v, dict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)):
        return type(v)(_wrap(x) for x in v)
    return v


def _merge(child, base):
    out = dict(base)
    for k, v in child.items():
        if isinstance(v, dict) and k in out and not v.get(DELETE_KEY, False):
            if not isinstance(out[k], dict):
                raise TypeError(f'{k}={v} in child config cannot inherit from base because {k} is a dict in the child '
                                f'config but is of type {type(out[k])} in base config. You may set `{DELETE_KEY}=True` '
                                'to ignore the base config')
            out[k] = _merge(v, out[k])
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != DELETE_KEY}
            out[k] = v
    return out


def _file2dict(filename):
    filename = os.path.abspath(os.path.expanduser(filename))
    if not os.path.isfile(filename):
        raise FileNotFoundError(f'file "{filename}" does not exist')
    if not filename.endswith('.py'):
        raise IOError('Only py type configs are supported')
    ns = runpy.run_path(filename)
    cfg = {k: v for k, v in ns.items() if not k.startswith('__') and not callable(v) and not hasattr(v, '__spec__')}
    if BASE_KEY in cfg:
        bases = cfg.pop(BASE_KEY)
        bases = bases if isinstance(bases, list) else [bases]
        merged = {}
        for b in bases:
            bd = _file2dict(os.path.join(os.path.dirname(filename), b))
            dup = merged.keys() & bd.keys()
            if dup:
                raise KeyError(f'Duplicate key is not allowed among bases: {sorted(dup)}')
            merged.update(bd)
        cfg = _merge(cfg, merged)
    return cfg


class Config:

    def __init__(self, cfg_dict=None, filename=None):
        if cfg_dict is None:
            cfg_dict = {}
        if not isinstance(cfg_dict, dict):
            raise TypeError(f'cfg_dict must be a dict, but got {type(cfg_dict)}')
        object.__setattr__(self, '_cfg', _wrap(cfg_dict))
        object.__setattr__(self, '_filename', filename)

    @staticmethod
    def fromfile(filename):
        return Config(_file2dict(filename), filename)

    filename = property(lambda self: self._filename)

    def __getattr__(self, name):
        return getattr(self._cfg, name)

    def __getitem__(self, name):
        return self._cfg[name]

    def __setattr__(self, name, value):
        self._cfg[name] = _wrap(value)

    def __contains__(self, name):
        return name in self._cfg

    def get(self, name, default=None):
        return self._cfg.get(name, default)

    def to_dict(self):
        return self._cfg
