"""ResNet / ResNeXt trunks with the reference's constructor arguments and parameter names
(mmdet/models/backbones/resnet.py:305-646, resnext.py:9-131).  The plain convolutions stay on cuDNN (bf16,
channels_last) in this round — SURVEY.md §8 row f4 ("next"); the DCNv2 ``conv2`` sites (stage_with_dcn) are built
through CONV_LAYERS and so land on the B200 deformable kernels (groups == 1 only for now)."""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.utils.checkpoint as cp
from torch.autograd import Function

from ..registry import BACKBONES, build_conv_layer

# ---------------------------------------------------------------------------------------------------------------
# Frozen-statistics BN folded into the convolution (mmdet/models/backbones/resnet.py:636-646 keeps every BN in eval
# mode; its affine parameters stay trainable).  With fixed (mean, var):  BN(conv(x, W)) = conv(x, W * s) + (beta - mean*s),
# s = gamma / sqrt(var + eps) -- exactly the same function of (W, gamma, beta), so autograd through the tiny folding
# ops yields the reference's gradients, while the three activation-sized BN passes (forward transform, backward
# reduce, backward element-wise) and the separate ReLU / residual-add passes disappear into cuDNN's fused
# conv+bias(+add)+ReLU epilogue.  Backward: one `grad_prep` pass (ReLU mask + bias column-sum) then cuDNN dgrad/wgrad.
# ---------------------------------------------------------------------------------------------------------------
_FUSED_OK = {'checked': False, 'ok': False}


def _fused_available(x):
    if not _FUSED_OK['checked']:
        _FUSED_OK['checked'] = True
        try:
            xx = torch.randn(1, 8, 8, 8, device=x.device, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
            ww = torch.randn(8, 8, 3, 3, device=x.device, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
            bb = torch.zeros(8, device=x.device, dtype=torch.bfloat16)
            y = torch.ops.aten.cudnn_convolution_relu(xx, ww, bb, [1, 1], [1, 1], [1, 1], 1)
            z = torch.ops.aten.cudnn_convolution_add_relu(xx, ww, y, 1.0, bb, [1, 1], [1, 1], [1, 1], 1)
            ref = F.relu(F.conv2d(xx.float(), ww.float(), None, 1, 1))
            _FUSED_OK['ok'] = bool(torch.isfinite(z).all()) and float((y.float() - ref).abs().max()) < 0.1
        except Exception:
            _FUSED_OK['ok'] = False
    return _FUSED_OK['ok']


class _ConvBiasAct(Function):
    """y = relu(conv(x, w) + bias (+ z)) on cuDNN's fused epilogue; bf16 channels_last."""

    @staticmethod
    def forward(ctx, x, w, bias, z, stride, padding, dilation, groups):
        from ..ops import gemm_ops as G
        x = G.as_nhwc(x, torch.bfloat16) if x.shape[1] >= 8 else x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wb = w.detach()
        if wb.dtype != torch.bfloat16 or not wb.is_contiguous(memory_format=torch.channels_last):
            wb = wb.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        bb = bias.detach().to(torch.bfloat16)
        if z is None:
            y = torch.ops.aten.cudnn_convolution_relu(x, wb, bb, stride, padding, dilation, groups)
        else:
            zz = z.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            y = torch.ops.aten.cudnn_convolution_add_relu(x, wb, zz, 1.0, bb, stride, padding, dilation, groups)
        y = G.as_nhwc(y)
        ctx.save_for_backward(x, wb, y)
        ctx.cfg = (stride, padding, dilation, groups, z is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        from ..ops import gemm_ops as G
        x, wb, y = ctx.saved_tensors
        stride, padding, dilation, groups, has_z = ctx.cfg
        g, colsum = G.grad_prep(gy, y, True, mult=1)          # ReLU mask + bias gradient in one pass
        gx, gw, _ = torch.ops.aten.convolution_backward(g, x, wb, None, stride, padding, dilation, False, [0, 0], groups,
                                                        [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        return gx, gw, colsum, (g if has_z else None), None, None, None, None


class _BnFold(Function):
    """(W, gamma, beta; frozen mean, var) -> (W' bf16 channels_last, b' fp32) in one kernel each way
    (lsnet_bn_fold_fwd / lsnet_bn_fold_bwd)."""

    @staticmethod
    def forward(ctx, W, gamma, beta, mean, var, eps):
        from .. import lib as L
        O, I, kh, kw = W.shape
        Wc = W.detach().contiguous()
        wb = torch.empty((O, kh, kw, I), device=W.device, dtype=torch.bfloat16)
        bias = torch.empty(O, device=W.device, dtype=torch.float32)
        L.call('lsnet_bn_fold_fwd', L.ptr(Wc), L.ptr(gamma.detach()), L.ptr(beta.detach()), L.ptr(mean), L.ptr(var),
               L.c_f(eps), L.c_int(O), L.c_int(I), L.c_int(kh * kw), L.ptr(wb), L.ptr(bias), L.stream())
        ctx.save_for_backward(Wc, gamma, mean, var)
        ctx.eps = eps
        return wb.permute(0, 3, 1, 2), bias       # logical OIHW, channels_last memory

    @staticmethod
    def backward(ctx, gwb, gbias):
        from .. import lib as L
        Wc, gamma, mean, var = ctx.saved_tensors
        O, I, kh, kw = Wc.shape
        gwb = gwb.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        gb = None if gbias is None else gbias.float().contiguous()
        gW = torch.empty_like(Wc)
        gg = torch.empty(O, device=Wc.device, dtype=torch.float32)
        gbt = torch.empty(O, device=Wc.device, dtype=torch.float32)
        L.call('lsnet_bn_fold_bwd', L.ptr(gwb), L.ptr(gb), L.ptr(Wc), L.ptr(gamma.detach()), L.ptr(mean), L.ptr(var),
               L.c_f(ctx.eps), L.c_int(O), L.c_int(I), L.c_int(kh * kw), L.ptr(gW), L.ptr(gg), L.ptr(gbt), L.stream())
        return gW, gg, gbt, None, None, None


def conv_bn_fold(conv, bn):
    """(W*s bf16 channels_last, beta - mean*s): folded weight and bias of a conv followed by an eval-mode BN."""
    return _BnFold.apply(conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, float(bn.eps))


class _BnFoldPacked(Function):
    """(W, gamma, beta; frozen mean, var) -> the two bf16 GEMM operand packs of W*s ([O, taps*I] and [I, taps*O]) and the
    folded bias, one kernel each way (lsnet_bn_fold2_fwd / _bwd).  The backward takes the weight gradient as the fp32
    [O, taps*I] matrix the weight-gradient GEMM wrote and -- when the trainer exposes the parameters' gradient memory --
    adds gW / ggamma / gbeta there directly."""

    @staticmethod
    def forward(ctx, W, gamma, beta, mean, var, eps, holder):
        # ``holder``: dict shared with the consuming _ConvPacked, which leaves the fp32 weight gradient there (autograd
        # would cast a gradient returned for the bf16 pack to bf16)
        from .. import lib as L
        ctx.holder = holder
        ctx.set_materialize_grads(False)     # no zero tensors for the packs' (absent) gradients
        O, I, kh, kw = W.shape
        KK = kh * kw
        Wd = W.detach()
        tap_major = KK > 1 and Wd.stride(1) == 1 and Wd.stride(3) == I and Wd.stride(2) == kw * I
        if not tap_major and not Wd.is_contiguous():
            Wd = Wd.contiguous()
        si, sk = (1, I) if tap_major else (KK, 1)
        wb = torch.empty((O, KK * I), device=W.device, dtype=torch.bfloat16)
        wt = torch.empty((I, KK * O), device=W.device, dtype=torch.bfloat16)
        bias = torch.empty(O, device=W.device, dtype=torch.float32)
        L.call('lsnet_bn_fold2_fwd', L.ptr(Wd), L.c_ll(si), L.c_ll(sk), L.ptr(gamma.detach()), L.ptr(beta.detach()),
               L.ptr(mean), L.ptr(var), L.c_f(eps), L.c_int(O), L.c_int(I), L.c_int(KK), L.ptr(wb), L.ptr(wt),
               L.ptr(bias), L.stream())
        ctx.save_for_backward(Wd, gamma, mean, var)
        ctx.cfg = (eps, si, sk, tap_major)
        ctx.params = (W, gamma, beta)
        ctx.mark_non_differentiable(wt)
        return wb, wt, bias

    @staticmethod
    def backward(ctx, gwb, _gwt, gbias):
        from .. import lib as L
        from ..ops import gemm_ops as G
        Wd, gamma, mean, var = ctx.saved_tensors
        eps, si, sk, tap_major = ctx.cfg
        O, I, kh, kw = Wd.shape
        KK = kh * kw
        gwb = ctx.holder.pop('gwb', None)
        if gwb is None:
            gwb = torch.zeros((O, KK * I), device=Wd.device, dtype=torch.float32)
        gb = None if gbias is None else gbias.float().contiguous()
        W, gm, bt = ctx.params
        # straight into the parameters' gradient memory (GraphTrainer's flat buffer) when all three are exposed
        tg, tb = G.direct_vec(gm), G.direct_vec(bt)
        tw = None
        if getattr(W, '_lsnet_direct_any', False) and W.grad is not None and W.grad.dtype == torch.float32 \
                and W.grad.stride() == Wd.stride():
            tw = W.grad
        direct = tw is not None and tg is not None and tb is not None and all(ctx.needs_input_grad[:3])
        if direct:
            gW, gg, gbt = tw, tg, tb
        else:
            gW = torch.empty_strided(Wd.shape, Wd.stride(), device=Wd.device, dtype=torch.float32)
            gg = torch.empty(O, device=Wd.device, dtype=torch.float32)
            gbt = torch.empty(O, device=Wd.device, dtype=torch.float32)
        L.call('lsnet_bn_fold2_bwd', L.ptr(gwb), L.ptr(gb), L.ptr(Wd), L.c_ll(si), L.c_ll(sk), L.ptr(gamma.detach()),
               L.ptr(mean), L.ptr(var), L.c_f(eps), L.c_int(O), L.c_int(I), L.c_int(KK), L.ptr(gW), L.ptr(gg),
               L.ptr(gbt), L.c_int(int(direct)), L.stream())
        if direct:
            return None, None, None, None, None, None, None
        return gW, gg, gbt, None, None, None, None


# 'own': trunk convolutions on the library's tcgen05 implicit-GEMM kernels (every groups == 1 conv whose input channels are
# a multiple of 64: all of ResNet-50/101 but the 3-channel stem); 'cudnn': the cuDNN fused conv+bias(+add)+ReLU path.
TRUNK = os.environ.get('LSNET_TRUNK', 'own')
# fold every (conv, BN) pair of the trunk on a side stream at the start of the forward (they only depend on parameters):
# 53 five-microsecond kernels leave the critical path, and so do their backward twins
FOLD_SIDE = os.environ.get('LSNET_FOLD_SIDE', '1') == '1'
STEM_OWN = os.environ.get('LSNET_STEM_OWN', '1') == '1'
# ReLU masks / identity-path sums of the trunk's backward inside the input-gradient GEMM epilogues (ops/conv.py::_ConvPacked)
BWD_FUSE = os.environ.get('LSNET_TRUNK_BWD_FUSE', '1') == '1'
_FOLD_STREAM = {}
_PREFOLD = {}


def _own_ok(conv):
    return (TRUNK == 'own' and isinstance(conv, nn.Conv2d) and conv.groups == 1 and conv.in_channels % 64 == 0
            and conv.out_channels % 16 == 0 and conv.kernel_size[0] == conv.kernel_size[1]
            and conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1]
            and conv.dilation[0] == conv.dilation[1] and conv.bias is None)


def _fold_apply(conv, bn):
    holder = {}
    return _BnFoldPacked.apply(conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, float(bn.eps),
                               holder) + (holder,)


def fold_packed(conv, bn):
    """(wb, wt, bias, holder) of a (conv, frozen BN) pair: from the side-stream prefold of this forward if there is one."""
    hit = _PREFOLD.get(id(conv))
    if hit is not None and hit[0] is conv:
        return hit[1]
    return _fold_apply(conv, bn)


def join_fold_stream(device):
    """After backward inside a stream capture: the fold backward nodes ran on the side stream -- rejoin it."""
    side = _FOLD_STREAM.get(str(device))
    if side is None:
        return
    with torch.cuda.stream(side):
        busy = torch.cuda.is_current_stream_capturing()
    if busy or not torch.cuda.is_current_stream_capturing():
        torch.cuda.current_stream(device).wait_stream(side)


def prefold(pairs, device):
    """Run the folds of ``pairs`` [(conv, bn)] on a side stream; returns the stream the consumers wait on."""
    cur = torch.cuda.current_stream(device)
    key = str(device)
    if key not in _FOLD_STREAM:
        _FOLD_STREAM[key] = torch.cuda.Stream(device=device)
    side = _FOLD_STREAM[key]
    side.wait_stream(cur)
    _PREFOLD.clear()
    with torch.cuda.stream(side):
        for conv, bn in pairs:
            out = _fold_apply(conv, bn)
            for t in out[:3]:
                t.record_stream(cur)
            _PREFOLD[id(conv)] = (conv, out)
    return side


def conv_bn_act(x, conv, bn, z=None, relu=True, extra_bias=None, opts=None):
    """relu(BN_eval(conv(x)) (+ z)) with the BN folded into the conv (see above)."""
    if _own_ok(conv):
        from ..ops.conv import conv2d_packed
        wb, wt, shift, holder = fold_packed(conv, bn)
        if extra_bias is not None:
            shift = shift + extra_bias
        return conv2d_packed(x, wb, wt, shift, z, conv.kernel_size, conv.stride[0], conv.padding[0], conv.dilation[0], relu,
                             wgrad_holder=holder, opts=opts if BWD_FUSE else None)
    w, shift = conv_bn_fold(conv, bn)
    if extra_bias is not None:
        shift = shift + extra_bias
    assert relu
    return _ConvBiasAct.apply(x, w, shift, z, list(conv.stride), list(conv.padding), list(conv.dilation), conv.groups)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style='pytorch', with_cp=False,
                 dcn=None, groups=1, base_width=4, base_channels=64, ref_parent_draws=False):
        super().__init__()
        assert style in ('pytorch', 'caffe')
        width = planes if groups == 1 else int(planes * (base_width / base_channels)) * groups   # resnext.py:33-37
        s1, s2 = (1, stride) if style == 'pytorch' else (stride, 1)
        self.with_cp, self.with_dcn = with_cp, dcn is not None
        if ref_parent_draws and os.environ.get('LSNET_REF_INIT_STREAM', '0') == '1':
            # The reference's ResNeXt block first runs the plain ResNet Bottleneck constructor -- planes-wide convs, its
            # own DCN pack -- and then replaces the three convs (resnext.py:14-24, 39-86).  Opt-in: draw the same random
            # numbers, so that a seeded build reproduces the reference's constructor-initialised DCN weights bit for bit.
            nn.Conv2d(inplanes, planes, 1, stride=s1, bias=False)
            d0 = dict(dcn) if dcn is not None else None
            if d0 is None or d0.pop('fallback_on_stride', False):
                nn.Conv2d(planes, planes, 3, stride=s2, padding=dilation, dilation=dilation, bias=False)
            else:
                build_conv_layer(d0, planes, planes, kernel_size=3, stride=s2, padding=dilation, dilation=dilation,
                                 bias=False)
            nn.Conv2d(planes, planes * self.expansion, 1, bias=False)
        self.conv1 = nn.Conv2d(inplanes, width, 1, stride=s1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        fallback = False
        if dcn is not None:
            dcn = dict(dcn)
            fallback = dcn.pop('fallback_on_stride', False)
        if dcn is None or fallback:
            self.conv2 = nn.Conv2d(width, width, 3, stride=s2, padding=dilation, dilation=dilation, groups=groups,
                                   bias=False)
        else:
            self.conv2 = build_conv_layer(dcn, width, width, kernel_size=3, stride=s2, padding=dilation,
                                          dilation=dilation, groups=groups, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, planes * self.expansion, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        for c in (self.conv1, self.conv2, self.conv3, None if downsample is None else downsample[0]):
            if c is not None and _own_ok(c):
                c.weight._lsnet_tapmajor = True      # GraphTrainer may keep it tap-major (see train.py)

    norm3 = property(lambda self: self.bn3)

    def _all_own(self):
        cs = [self.conv1, self.conv2, self.conv3] + ([] if self.downsample is None else [self.downsample[0]])
        return all(_own_ok(c) for c in cs)

    def _fusable(self, x):
        # (the cuDNN fused conv+bias+ReLU probe only matters for convs that are not on the library's kernels)
        return (x.is_cuda and not self.bn1.training and not self.bn2.training and not self.bn3.training
                and isinstance(self.conv2, nn.Conv2d) and (self._all_own() or _fused_available(x)))

    def _inner_fused(self, x):
        own = _own_ok(self.conv1) and _own_ok(self.conv2) and _own_ok(self.conv3) and \
            (self.downsample is None or _own_ok(self.downsample[0]))
        if own:
            # Backward fusion (exact, see _ConvPacked): inside the block every ReLU output has ONE consumer, so the
            # consumer's input-gradient GEMM applies the mask and the producer only sums its bias columns; the block input
            # is read by conv1 and the identity path (or the downsample conv) -- they share a gradient sink.
            # in_relu: x is the (ReLU) output of the previous Bottleneck; out_premasked: every consumer of this block's
            # output masks by it (the next Bottleneck of the stage) -- both set by ResNet.
            # the sink needs EVERY consumer of x inside this block (autograd may sum an outside gradient with the first
            # one before the second is added in place): true for all but a stage's first block, whose input also feeds
            # the neck / is the previous stage's output
            sink = {} if getattr(self, 'in_exclusive', False) else None
            in_mask = getattr(self, 'in_relu', False)
            out = conv_bn_act(x, self.conv1, self.bn1, opts=dict(mask_input=in_mask, premasked=True, x_sink=sink))
            out = conv_bn_act(out, self.conv2, self.bn2, opts=dict(mask_input=True, premasked=True))
            # (_rt_premasked: ResNet.forward re-checks the promise at run time -- it only holds while the NEXT block also
            # takes this fused path)
            o3 = dict(mask_input=True, premasked=bool(getattr(self, 'out_premasked', False) and
                                                      getattr(self, '_rt_premasked', False)))
            if self.downsample is None:
                return conv_bn_act(out, self.conv3, self.bn3, z=x, opts=dict(o3, z_sink=sink))
            ident = conv_bn_act(x, self.downsample[0], self.downsample[1], relu=False,
                                opts=dict(mask_input=in_mask, x_sink=sink))
            return conv_bn_act(out, self.conv3, self.bn3, z=ident, opts=o3)
        out = conv_bn_act(x, self.conv1, self.bn1)
        out = conv_bn_act(out, self.conv2, self.bn2)
        if self.downsample is None:
            return conv_bn_act(out, self.conv3, self.bn3, z=x)
        dc = self.downsample[0]
        if _own_ok(dc) and _own_ok(self.conv3):
            # identity branch on the same kernels (bias in its own epilogue, no ReLU), added in conv3's epilogue
            ident = conv_bn_act(x, dc, self.downsample[1], relu=False)
            return conv_bn_act(out, self.conv3, self.bn3, z=ident)
        # identity branch: plain conv with the folded weight; its folded bias rides on conv3's bias
        wd, bd = conv_bn_fold(dc, self.downsample[1])
        ident = F.conv2d(x.to(torch.bfloat16), wd, None, dc.stride, dc.padding)
        w3, b3 = conv_bn_fold(self.conv3, self.bn3)
        return _ConvBiasAct.apply(out, w3, b3 + bd, ident, [1, 1], [0, 0], [1, 1], 1)

    def _inner(self, x):
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return out + (x if self.downsample is None else self.downsample(x))

    def forward(self, x):
        if self._fusable(x):       # ReLU of the block output is inside the fused epilogue
            if self.with_cp and x.requires_grad:
                return cp.checkpoint(self._inner_fused, x, use_reentrant=False)
            return self._inner_fused(x)
        out = cp.checkpoint(self._inner, x, use_reentrant=False) if self.with_cp and x.requires_grad else self._inner(x)
        return self.relu(out)


@BACKBONES.register_module()
class ResNet(nn.Module):
    arch_settings = {50: (Bottleneck, (3, 4, 6, 3)), 101: (Bottleneck, (3, 4, 23, 3)), 152: (Bottleneck, (3, 8, 36, 3))}
    _resnext_blocks = False      # ResNeXt: the reference's block class constructs its convs twice (see Bottleneck)

    def __init__(self, depth, in_channels=3, stem_channels=64, base_channels=64, num_stages=4, strides=(1, 2, 2, 2),
                 dilations=(1, 1, 1, 1), out_indices=(0, 1, 2, 3), style='pytorch', deep_stem=False, avg_down=False,
                 frozen_stages=-1, conv_cfg=None, norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True,
                 dcn=None, stage_with_dcn=(False, False, False, False), plugins=None, with_cp=False,
                 zero_init_residual=True, groups=1, base_width=4):
        super().__init__()
        if depth not in self.arch_settings:
            raise KeyError(f'invalid depth {depth} for resnet')
        if deep_stem or avg_down or plugins is not None or conv_cfg is not None:
            raise NotImplementedError('deep_stem / avg_down / plugins / conv_cfg are not on the LSNet path')
        assert norm_cfg.get('type', 'BN') == 'BN'
        # resnet.py:377-391: the constructor's own argument checks
        assert 1 <= num_stages <= 4
        assert len(strides) == len(dilations) == num_stages
        assert max(out_indices) < num_stages
        assert style in ('pytorch', 'caffe')
        if dcn is not None:
            assert len(stage_with_dcn) == num_stages
        block, stage_blocks = self.arch_settings[depth]
        self.depth, self.out_indices, self.frozen_stages = depth, out_indices, frozen_stages
        self.norm_eval, self.zero_init_residual, self.dcn = norm_eval, zero_init_residual, dcn
        self.norm_requires_grad = norm_cfg.get('requires_grad', True)
        self.conv1 = nn.Conv2d(in_channels, stem_channels, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(stem_channels)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        self.res_layers = []
        inplanes = stem_channels
        for i, nblocks in enumerate(stage_blocks[:num_stages]):
            planes = base_channels * 2 ** i
            layers = []
            for j in range(nblocks):
                stride = strides[i] if j == 0 else 1
                down = None
                if j == 0 and (stride != 1 or inplanes != planes * block.expansion):
                    down = nn.Sequential(nn.Conv2d(inplanes, planes * block.expansion, 1, stride=stride, bias=False),
                                         nn.BatchNorm2d(planes * block.expansion))
                layers.append(block(inplanes, planes, stride, dilations[i], down, style, with_cp,
                                    dcn if stage_with_dcn[i] else None, groups, base_width, base_channels,
                                    ref_parent_draws=self._resnext_blocks))
                inplanes = planes * block.expansion
            # backward-fusion topology (Bottleneck._inner_fused): a block's input is the previous block's ReLU output
            # (the first block of stage 1 reads the max-pool instead); the output of every block but the stage's last has
            # the next block as its only consumer (stage outputs also feed the neck)
            def all_own(b):
                cs = [b.conv1, b.conv2, b.conv3] + ([] if b.downsample is None else [b.downsample[0]])
                return all(_own_ok(c) for c in cs)
            for j, blk in enumerate(layers):
                blk.in_relu = not (i == 0 and j == 0)
                blk.in_exclusive = j > 0
                blk.out_premasked = j + 1 < nblocks and not with_cp and all_own(layers[j + 1])
            name = f'layer{i + 1}'
            self.add_module(name, nn.Sequential(*layers))
            self.res_layers.append(name)
        if not self.norm_requires_grad:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    for p in m.parameters():
                        p.requires_grad = False
        self._freeze_stages()
        self.feat_dim = inplanes

    norm1 = property(lambda self: self.bn1)      # resnet.py:467-470 (the norm layer is registered as 'bn1')

    def _freeze_stages(self):          # resnet.py:569-585
        if self.frozen_stages >= 0:
            self.bn1.eval()
            for m in (self.conv1, self.bn1):
                for p in m.parameters():
                    p.requires_grad = False
        for i in range(1, self.frozen_stages + 1):
            m = getattr(self, f'layer{i}')
            m.eval()
            for p in m.parameters():
                p.requires_grad = False

    def init_weights(self, pretrained=None):   # resnet.py:587-617
        if isinstance(pretrained, str):
            sd = torch.load(pretrained, map_location='cpu')
            self.load_state_dict(sd.get('state_dict', sd), strict=False)
            return
        if pretrained is not None:
            raise TypeError('pretrained must be a str or None')
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        for m in self.modules():
            if isinstance(m, Bottleneck):
                if hasattr(m.conv2, 'conv_offset'):
                    nn.init.constant_(m.conv2.conv_offset.weight, 0)
                    nn.init.constant_(m.conv2.conv_offset.bias, 0)
                if self.zero_init_residual:
                    nn.init.constant_(m.bn3.weight, 0)

    def _own_stem(self, x):
        """The frozen stem (frozen_stages >= 0: no gradient reaches conv1 / bn1) on the library's stem kernels."""
        c, m = self.conv1, self.maxpool
        return (TRUNK == 'own' and STEM_OWN and not x.requires_grad and not c.weight.requires_grad
                and not self.bn1.weight.requires_grad and c.in_channels == 3 and c.out_channels == 64
                and c.kernel_size == (7, 7) and c.stride == (2, 2) and c.padding == (3, 3) and c.bias is None
                and m.kernel_size == 3 and m.stride == 2 and m.padding == 1 and x.dtype in (torch.float32, torch.bfloat16))

    def _fold_pairs(self):
        pairs = []
        for m in self.modules():
            if isinstance(m, Bottleneck) and not m.bn1.training:
                cands = [(m.conv1, m.bn1), (m.conv2, m.bn2), (m.conv3, m.bn3)]
                if m.downsample is not None:
                    cands.append((m.downsample[0], m.downsample[1]))
                pairs += [(c, b) for c, b in cands if _own_ok(c)]
        return pairs

    def forward(self, x):
        fold_ev = None
        if x.is_cuda and not self.bn1.training and (self._own_stem(x) or _fused_available(x)):
            if FOLD_SIDE and TRUNK == 'own':
                fold_side = prefold(self._fold_pairs(), x.device)
                fold_ev = True
            if self._own_stem(x):
                from .. import ops
                wp, shift = ops.gemm_ops.cached_pack(
                    self.conv1.weight, 'stem', lambda t: ops.pack_stem_weight(
                        t, self.bn1.weight.detach(), self.bn1.bias.detach(), self.bn1.running_mean, self.bn1.running_var,
                        self.bn1.eps))
                x = ops.maxpool3x3s2(ops.stem_conv(x, wp, shift))
            else:
                x = self.maxpool(conv_bn_act(x, self.conv1, self.bn1))
            if fold_ev:
                torch.cuda.current_stream(x.device).wait_stream(fold_side)
        else:
            x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        outs = []
        for name in self.res_layers:      # run-time half of the backward-fusion topology (see __init__)
            blocks = list(getattr(self, name))
            for j, blk in enumerate(blocks):
                blk._rt_premasked = j + 1 < len(blocks) and blocks[j + 1]._fusable(x) and not blocks[j + 1].with_cp
        for i, name in enumerate(self.res_layers):
            x = getattr(self, name)(x)
            if i in self.out_indices:
                outs.append(x)
        _PREFOLD.clear()
        return tuple(outs)

    def train(self, mode=True):        # resnet.py:636-646
        super().train(mode)
        self._freeze_stages()
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.eval()
        return self


@BACKBONES.register_module()
class ResNeXt(ResNet):
    """resnext.py:76-131: Bottleneck width = planes * base_width / 64 * groups."""
    _resnext_blocks = True

    def __init__(self, groups=1, base_width=4, **kwargs):
        super().__init__(groups=groups, base_width=base_width, **kwargs)
