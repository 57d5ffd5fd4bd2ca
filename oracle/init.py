"""ORACLE — test infrastructure only.  Deterministic, seed-reproducible LSNet state_dicts with the reference's
parameter names and shapes, so that golden vectors can be stored without the 150 MB of
weights: the generating script loads ``make_state_dict(task, seed)`` into the *reference* modules, and the tests
rebuild the identical tensors from the seed (torch's CPU generator is platform-stable).

Values are NOT the reference's init rules (e.g. conv_offset is non-zero here so that the DCN sampling path is
exercised); the reference's rules live in the product modules' ``init_weights``.
"""
import math

import torch

NUM_VECTORS = {'bbox': 4, 'segm': 36, 'pose_bbox': 17}


def _conv(sd, g, name, cout, cin, k, bias, std=None):
    fan_in = cin * k * k
    std = std if std is not None else math.sqrt(2.0 / fan_in)
    sd[name + '.weight'] = torch.randn(cout, cin, k, k, generator=g) * std
    if bias:
        sd[name + '.bias'] = torch.randn(cout, generator=g) * 0.01


def _norm(sd, g, name, c, bn=False):
    sd[name + '.weight'] = 1.0 + 0.1 * torch.randn(c, generator=g)
    sd[name + '.bias'] = 0.05 * torch.randn(c, generator=g)
    if bn:
        sd[name + '.running_mean'] = 0.05 * torch.randn(c, generator=g)
        sd[name + '.running_var'] = 1.0 + 0.1 * torch.rand(c, generator=g)
        sd[name + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)


def resnet50_state(sd, g, prefix='backbone.'):
    _conv(sd, g, prefix + 'conv1', 64, 3, 7, False)
    _norm(sd, g, prefix + 'bn1', 64, bn=True)
    inplanes = 64
    for li, (planes, blocks) in enumerate(zip((64, 128, 256, 512), (3, 4, 6, 3)), 1):
        for bi in range(blocks):
            p = f'{prefix}layer{li}.{bi}.'
            _conv(sd, g, p + 'conv1', planes, inplanes, 1, False)
            _norm(sd, g, p + 'bn1', planes, bn=True)
            _conv(sd, g, p + 'conv2', planes, planes, 3, False)
            _norm(sd, g, p + 'bn2', planes, bn=True)
            _conv(sd, g, p + 'conv3', planes * 4, planes, 1, False)
            _norm(sd, g, p + 'bn3', planes * 4, bn=True)
            sd[p + 'bn3.weight'] = sd[p + 'bn3.weight'] * 0.25      # keep the residual stack well-scaled
            if bi == 0:
                _conv(sd, g, p + 'downsample.0', planes * 4, inplanes, 1, False)
                _norm(sd, g, p + 'downsample.1', planes * 4, bn=True)
            inplanes = planes * 4


def fpn_state(sd, g, prefix='neck.', in_channels=(512, 1024, 2048), out=256):
    for i, c in enumerate(in_channels):
        _conv(sd, g, f'{prefix}lateral_convs.{i}.conv', out, c, 1, False)
        _norm(sd, g, f'{prefix}lateral_convs.{i}.gn', out)
    for i in range(5):
        cin = in_channels[-1] if i == 3 else out
        _conv(sd, g, f'{prefix}fpn_convs.{i}.conv', out, cin, 3, False)
        _norm(sd, g, f'{prefix}fpn_convs.{i}.gn', out)


def head_state(sd, g, task='bbox', prefix='bbox_head.', c=256, num_classes=None):
    nv = NUM_VECTORS[task]
    num_classes = num_classes if num_classes is not None else (1 if task == 'pose_bbox' else 80)
    branches = {'bbox': ['bbox'], 'segm': ['segm'], 'pose_bbox': ['bbox', 'pose']}[task]
    _norm(sd, g, prefix + 'cls_GN', c)
    towers = ['cls'] + branches
    for t in towers:
        if t != 'cls':
            _norm(sd, g, f'{prefix}{t}_GN', c)
        for i in range(3):
            p = f'{prefix}{t}_convs.{i}.'
            _conv(sd, g, p + 'conv', c, c, 3, True)
            _conv(sd, g, p + 'conv.conv_offset', 27, c, 3, True, std=0.02)
            _norm(sd, g, p + 'bn', c)
    _conv(sd, g, prefix + 'pts_cls_conv', c, c, 3, False)
    _conv(sd, g, prefix + 'pts_cls_out', num_classes, c, 1, True, std=0.01)
    sd[prefix + 'pts_cls_out.bias'] = sd[prefix + 'pts_cls_out.bias'] - 4.595
    _conv(sd, g, prefix + 'cls_af_dcn_conv.0', c, 3 * c, 1, True)
    _conv(sd, g, prefix + 'cls_feat_conv', c, c, 3, True)
    for br in branches:
        if br == 'bbox':
            init_dim, ref_dim = 28, 20
        else:
            init_dim = ref_dim = 4 * (nv + 1)
        _conv(sd, g, f'{prefix}pts_{br}_init_conv', c, c, 3, True)
        _conv(sd, g, f'{prefix}pts_{br}_init_out', init_dim, c, 1, True, std=0.02)
        _conv(sd, g, f'{prefix}pts_{br}_refine_conv', c, c, 3, False)
        _conv(sd, g, f'{prefix}pts_{br}_refine_out', ref_dim, c, 1, True, std=0.02)
        _conv(sd, g, f'{prefix}{br}_af_dcn_conv.0', c, 3 * c, 1, True)
        _conv(sd, g, f'{prefix}{br}_feat_conv', c, c, 3, True)


def make_state_dict(task='bbox', seed=0, parts=('backbone', 'neck', 'head')):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    if 'backbone' in parts:
        resnet50_state(sd, g)
    if 'neck' in parts:
        fpn_state(sd, g)
    if 'head' in parts:
        head_state(sd, g, task)
    return sd
