"""nn.Module wrappers of the deformable convolutions with the reference's constructor arguments, parameter names
and registry names (mmdet/ops/dcn/deform_conv.py:295-630): ``DeformConv``, ``DeformConvPack`` ('DCN'),
``ModulatedDeformConv``, ``ModulatedDeformConvPack`` ('DCNv2'), ``PyramidDeformConv``."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.modules.utils import _pair, _single

from .. import ops
from ..registry import CONV_LAYERS


import os

GX_SINK = os.environ.get('LSNET_GX_SINK', '1') == '1'
PACKED_OFFSET_MASK = os.environ.get('LSNET_DCN_PACKED_OM', '1') == '1'


class _DeformBase(nn.Module):

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=False):
        super().__init__()
        assert in_channels % groups == 0, f'in_channels {in_channels} is not divisible by groups {groups}'
        assert out_channels % groups == 0, f'out_channels {out_channels} is not divisible by groups {groups}'
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride, self.padding, self.dilation = _pair(stride), _pair(padding), _pair(dilation)
        self.groups, self.deformable_groups = groups, deformable_groups
        self.transposed, self.output_padding = False, _single(0)   # nn.Conv2d compatibility (ConvModule copies them)
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        self.weight._lsnet_tapmajor = groups == 1     # GraphTrainer may keep it tap-major (see train.py)
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1. / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)


class DeformConv(_DeformBase):
    """deform_conv.py:295-357 (DCNv1, no bias)."""

    def __init__(self, *args, bias=False, **kwargs):
        assert not bias
        super().__init__(*args, **kwargs)

    def forward(self, x, offset):
        ph = max(self.kernel_size[0] - x.size(2), 0)
        pw = max(self.kernel_size[1] - x.size(3), 0)
        if ph or pw:   # deform_conv.py:341-349
            x = F.pad(x, (0, pw, 0, ph))
            offset = F.pad(offset, (0, pw, 0, ph))
        out = ops.deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups,
                              self.deformable_groups)
        if ph or pw:
            out = out[:, :, :out.size(2) - ph, :out.size(3) - pw]
        return out


@CONV_LAYERS.register_module('DCN')
class DeformConvPack(DeformConv):
    """deform_conv.py:360-435: offsets from a zero-initialised conv."""
    _version = 2

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_offset = nn.Conv2d(self.in_channels, self.deformable_groups * 2 * self.kernel_size[0] *
                                     self.kernel_size[1], kernel_size=self.kernel_size, stride=self.stride,
                                     padding=self.padding, dilation=self.dilation, bias=True)
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()

    def forward(self, x):
        offset = _offset_conv(self, x)
        return ops.deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups,
                               self.deformable_groups)


class ModulatedDeformConv(_DeformBase):
    """deform_conv.py:438-485 (DCNv2)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=True):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, deformable_groups)
        self.with_bias = bias
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter('bias', None)

    def forward(self, x, offset, mask):
        return ops.modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride, self.padding,
                                         self.dilation, self.groups, self.deformable_groups)


def _offset_conv(m, x, gx_sink=None):
    """conv_offset through the tcgen05 implicit-GEMM kernel when it is a stride-1 'same' conv, else cuDNN."""
    k, s, p, d = m.kernel_size, m.stride, m.padding, m.dilation
    if s == (1, 1) and k[0] == k[1] and p[0] == p[1] and d[0] == d[1] and 2 * p[0] == d[0] * (k[0] - 1) and x.is_cuda:
        return ops.conv2d_same(x, m.conv_offset.weight, m.conv_offset.bias, padding=p[0], dilation=d[0], out_fp32=True,
                               gx_sink=gx_sink)
    return m.conv_offset(x)


@CONV_LAYERS.register_module('DCNv2')
class ModulatedDeformConvPack(ModulatedDeformConv):
    """deform_conv.py:488-562: offsets + mask from one zero-initialised 3x3 conv (27 channels)."""
    _version = 2

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_offset = nn.Conv2d(self.in_channels, self.deformable_groups * 3 * self.kernel_size[0] *
                                     self.kernel_size[1], kernel_size=self.kernel_size, stride=self.stride,
                                     padding=self.padding, dilation=self.dilation, bias=True)
        self.conv_offset.weight._lsnet_tapmajor = self.stride == (1, 1)
        self.init_offset()
        if os.environ.get('LSNET_REF_INIT_STREAM', '0') == '1':
            # the reference draws this weight a second time AFTER conv_offset's default initialisation (the base class
            # constructor and the pack constructor both end in the overridden init_weights, deform_conv.py:470-472,
            # 519-525); trunk DCN weights keep their constructor value (ResNet.init_weights only re-draws nn.Conv2d), so
            # reproducing the reference's seeded initial model bit for bit needs the same second draw.  Opt-in: it shifts
            # the random stream of everything constructed afterwards.
            self.reset_parameters()

    def init_offset(self):
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()

    def forward(self, x, exclusive=False, gn_holder=None, gn_groups=0, skip_bias_grad=False):
        """``exclusive``: the caller guarantees that this module is the ONLY consumer of ``x`` in the autograd graph (a
        tower layer fed by the previous layer).  Then the two internal consumers of x share a gradient sink: the sampling
        op's backward runs first (conv_offset's output feeds it) and the conv_offset backward adds its input gradient into
        that tensor inside its GEMM epilogue.  With other consumers autograd may already have summed the first gradient
        into a new tensor when the second arrives, so the in-place sum would be lost -- hence the explicit promise."""
        n = self.deformable_groups * self.kernel_size[0] * self.kernel_size[1]
        if PACKED_OFFSET_MASK and x.is_cuda:
            # one op: the sampling kernels split offsets / mask logits and apply the sigmoid (and its derivative)
            sink = {} if (GX_SINK and exclusive and x.requires_grad and x.dtype == torch.bfloat16) else None
            out = _offset_conv(self, x, sink)
            # gn_holder / gn_groups: the GroupNorm that follows takes its statistics from this op's GEMM epilogue
            return ops.modulated_deform_conv_packed(x, out[:, :3 * n], self.weight, self.bias, self.stride, self.padding,
                                                    self.dilation, self.groups, self.deformable_groups, gx_sink=sink,
                                                    gn_holder=gn_holder, gn_groups=gn_groups,
                                                    skip_bias_grad=skip_bias_grad)
        out = _offset_conv(self, x)
        # chunk(3) + cat(o1, o2) keeps the channel order (deform_conv.py:528-531): offsets = first 2n channels
        offset, mask = out[:, :2 * n], torch.sigmoid(out[:, 2 * n:3 * n])
        return ops.modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride, self.padding,
                                         self.dilation, self.groups, self.deformable_groups)


class PyramidDeformConv(_DeformBase):
    """deform_conv.py:565-630: the sampling grid lives on the offset map's level, x is another FPN level."""

    def __init__(self, *args, bias=False, **kwargs):
        assert not bias
        super().__init__(*args, **kwargs)

    def forward(self, x, offset, scale_h, scale_w, out_slice=None):
        """``out_slice=(buffer [B,H,W,C_total] bf16, channel offset)`` makes the op write its result in place into a
        wider pixel-major buffer (LSHead concatenates three of these along channels)."""
        ph = max(self.kernel_size[0] - x.size(2), 0)
        pw = max(self.kernel_size[1] - x.size(3), 0)
        if ph or pw:
            x = F.pad(x, (0, pw, 0, ph))
            offset = F.pad(offset, (0, pw, 0, ph))
            out_slice = None
        out = ops.pyramid_deform_conv(x, offset, self.weight, (scale_h, scale_w), self.stride, self.padding,
                                      self.dilation, self.groups, self.deformable_groups, out_slice=out_slice)
        if ph or pw:
            out = out[:, :, :out.size(2) - ph, :out.size(3) - pw]
        return out
