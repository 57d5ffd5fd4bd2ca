"""FPN neck with the reference's constructor arguments and parameter names (mmdet/models/necks/fpn.py:65-217):
lateral 1x1 + norm, nearest top-down add, 3x3 output convs + norm, extra stride-2 levels.  The stride-1 convolutions
run on the tcgen05 kernels; the two stride-2 extra-level convs (13x21 / 7x11 maps at 800x1344) use cuDNN."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..registry import NECKS


class ConvNorm(nn.Module):
    """The subset of mmcv ConvModule FPN uses (conv -> norm, no activation; conv bias dropped when a norm follows,
    mmcv/mmcv/cnn/bricks/conv_module.py:91-93; the norm attribute is named by its abbreviation, :136-137)."""

    def __init__(self, cin, cout, k, stride=1, padding=0, norm_cfg=None, act=False):
        super().__init__()
        self.with_norm = norm_cfg is not None
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=padding, bias=not self.with_norm)
        self.stride, self.padding, self.k, self.act = stride, padding, k, act
        self.conv.weight._lsnet_tapmajor = (stride == 1 and (k == 1 or padding == k // 2)) or \
            (cin % 64 == 0 and cout % 64 == 0)   # see train.py
        if self.with_norm:
            kind = norm_cfg.get('type', 'GN')
            if kind == 'GN':
                self.gn = nn.GroupNorm(norm_cfg['num_groups'], cout)
            elif kind == 'BN':
                self.bn = nn.BatchNorm2d(cout)
            else:
                raise KeyError(f'Unrecognized norm type {kind}')
            if not norm_cfg.get('requires_grad', True):
                for p in self.norm.parameters():
                    p.requires_grad = False

    @property
    def norm(self):
        return getattr(self, 'gn', None) or getattr(self, 'bn', None)

    def forward(self, x):
        if self.stride == 1 and x.is_cuda and (self.k == 1 or self.padding == self.k // 2):
            x = ops.conv2d_same(x, self.conv.weight, self.conv.bias, padding=self.padding)
        elif x.is_cuda and self.conv.in_channels % 64 == 0 and self.conv.out_channels % 64 == 0 and self.conv.groups == 1:
            # the stride-2 extra levels (fpn.py:203-211): strided TMA boxes forward, one implicit GEMM per output phase backward
            x = ops.conv2d_strided(x, self.conv.weight, self.conv.bias, self.stride, self.padding)
        else:
            x = F.conv2d(x.to(torch.bfloat16), self.conv.weight.to(torch.bfloat16),
                         None if self.conv.bias is None else self.conv.bias.to(torch.bfloat16), self.stride, self.padding)
        if self.with_norm:
            n = self.norm
            if isinstance(n, nn.GroupNorm):
                return ops.group_norm_nhwc(x, n.num_groups, n.weight, n.bias, n.eps, relu=self.act)
            x = n(x)
        return F.relu(x) if self.act else x


@NECKS.register_module()
class FPN(nn.Module):

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, add_extra_convs=False,
                 extra_convs_on_inputs=True, relu_before_extra_convs=False, no_norm_on_lateral=False, conv_cfg=None,
                 norm_cfg=None, act_cfg=None, upsample_cfg=dict(mode='nearest')):
        super().__init__()
        assert isinstance(in_channels, list)
        self.in_channels, self.out_channels, self.num_outs = in_channels, out_channels, num_outs
        self.num_ins = len(in_channels)
        self.relu_before_extra_convs, self.upsample_cfg = relu_before_extra_convs, dict(upsample_cfg)
        if end_level == -1:
            self.backbone_end_level = self.num_ins
            assert num_outs >= self.num_ins - start_level
        else:
            self.backbone_end_level = end_level
            assert end_level <= len(in_channels) and num_outs == end_level - start_level
        self.start_level, self.end_level = start_level, end_level
        assert isinstance(add_extra_convs, (str, bool))
        if isinstance(add_extra_convs, str):
            assert add_extra_convs in ('on_input', 'on_lateral', 'on_output')
        elif add_extra_convs:
            add_extra_convs = 'on_input' if extra_convs_on_inputs else 'on_output'
        self.add_extra_convs = add_extra_convs
        self.lateral_convs, self.fpn_convs = nn.ModuleList(), nn.ModuleList()
        for i in range(self.start_level, self.backbone_end_level):
            self.lateral_convs.append(ConvNorm(in_channels[i], out_channels, 1,
                                               norm_cfg=None if no_norm_on_lateral else norm_cfg))
            self.fpn_convs.append(ConvNorm(out_channels, out_channels, 3, padding=1, norm_cfg=norm_cfg))
        extra = num_outs - self.backbone_end_level + self.start_level
        if self.add_extra_convs and extra >= 1:
            for i in range(extra):
                cin = self.in_channels[self.backbone_end_level - 1] if (i == 0 and self.add_extra_convs == 'on_input') \
                    else out_channels
                self.fpn_convs.append(ConvNorm(cin, out_channels, 3, stride=2, padding=1, norm_cfg=norm_cfg))

    def init_weights(self):            # fpn.py:159-163
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)
        lat = [lc(inputs[i + self.start_level]) for i, lc in enumerate(self.lateral_convs)]
        n = len(lat)
        for i in range(n - 1, 0, -1):
            if 'scale_factor' in self.upsample_cfg:
                lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], **self.upsample_cfg)
            elif lat[i].is_cuda and self.upsample_cfg.get('mode', 'nearest') == 'nearest' and lat[i].shape[1] % 8 == 0 \
                    and len(self.upsample_cfg) == 1:
                lat[i - 1] = ops.upsample_add(lat[i - 1], lat[i])       # one kernel each way instead of upsample + add
            else:
                lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], **self.upsample_cfg)
        outs = [self.fpn_convs[i](lat[i]) for i in range(n)]
        if self.num_outs > len(outs):
            if not self.add_extra_convs:
                for _ in range(self.num_outs - n):
                    outs.append(F.max_pool2d(outs[-1], 1, stride=2))
            else:
                src = {'on_input': inputs[self.backbone_end_level - 1], 'on_lateral': lat[-1],
                       'on_output': outs[-1]}[self.add_extra_convs]
                outs.append(self.fpn_convs[n](src))
                for i in range(n + 1, self.num_outs):
                    outs.append(self.fpn_convs[i](F.relu(outs[-1]) if self.relu_before_extra_convs else outs[-1]))
        return tuple(outs)
