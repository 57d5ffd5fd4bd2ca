"""ORACLE — test infrastructure only (see oracle/__init__.py).

CPU autograd ops for the reference's three deformable convolutions, built on the
C restatement in ``dcn_ref.c``.  The GEMM sequencing follows the reference host
code (mmdet/ops/dcn/src/cuda/deform_conv_cuda.cpp:274-1154): per group
``out_g = W_g . col_g (+ bias)``; backward ``gcol_g = W_g^T . dY_g`` ->
col2im_coord (dOffset, dMask) -> col2im (dX); ``dW_g = dY_g . col_g^T``;
``dbias = sum dY``.  Signatures mirror mmdet/ops/dcn/deform_conv.py:290-292
(``deform_conv``, ``modulated_deform_conv``, ``pyramid_deform_conv``).
"""
import ctypes
import os
import subprocess

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """gcc-compile dcn_ref.c -> oracle/_build/libdcn_ref.so (idempotent)."""
    src = os.path.join(_HERE, 'dcn_ref.c')
    out_dir = os.path.join(_HERE, '_build')
    out = os.path.join(out_dir, 'libdcn_ref.so')
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-fopenmp', '-shared', '-fPIC', '-o', out, src, '-lm'])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _suf(t):
    if t.dtype == torch.float32:
        return 'f32'
    if t.dtype == torch.float64:
        return 'f64'
    raise TypeError(f'oracle DCN supports fp32/fp64, got {t.dtype}')


def _geom(x, Ho, Wo, kh, kw, stride, pad, dil, scales, dg):
    B, C, H, W = x.shape
    return [ctypes.c_int(v) for v in (B, C, H, W, Ho, Wo, kh, kw, stride[0], stride[1], pad[0], pad[1],
                                      dil[0], dil[1])] + \
           [ctypes.c_float(scales[0]), ctypes.c_float(scales[1]), ctypes.c_int(dg)]


def im2col(x, offset, mask, Ho, Wo, kh, kw, stride, pad, dil, scales, dg):
    x, offset = x.contiguous(), offset.contiguous()
    mask = mask.contiguous() if mask is not None else None
    B, C = x.shape[:2]
    col = x.new_zeros((C * kh * kw, B, Ho, Wo))
    getattr(_lib(), 'dcn_im2col_' + _suf(x))(_p(x), _p(offset), _p(mask),
                                             *_geom(x, Ho, Wo, kh, kw, stride, pad, dil, scales, dg), _p(col))
    return col


def col2im(gcol, x_like, offset, mask, Ho, Wo, kh, kw, stride, pad, dil, scales, dg):
    gcol, offset = gcol.contiguous(), offset.contiguous()
    mask = mask.contiguous() if mask is not None else None
    gx = torch.zeros_like(x_like, memory_format=torch.contiguous_format)
    getattr(_lib(), 'dcn_col2im_' + _suf(gx))(_p(gcol), _p(offset), _p(mask),
                                              *_geom(gx, Ho, Wo, kh, kw, stride, pad, dil, scales, dg), _p(gx))
    return gx


def col2im_coord(gcol, x, offset, mask, Ho, Wo, kh, kw, stride, pad, dil, scales, dg):
    gcol, x, offset = gcol.contiguous(), x.contiguous(), offset.contiguous()
    mask = mask.contiguous() if mask is not None else None
    goff = torch.zeros_like(offset, memory_format=torch.contiguous_format)
    gmask = torch.zeros_like(mask, memory_format=torch.contiguous_format) if mask is not None else None
    getattr(_lib(), 'dcn_col2im_coord_' + _suf(x))(_p(gcol), _p(x), _p(offset), _p(mask),
                                                   *_geom(x, Ho, Wo, kh, kw, stride, pad, dil, scales, dg),
                                                   _p(goff), _p(gmask))
    return goff, gmask


def _out_hw(H, W, kh, kw, stride, pad, dil):
    Ho = (H + 2 * pad[0] - (dil[0] * (kh - 1) + 1)) // stride[0] + 1
    Wo = (W + 2 * pad[1] - (dil[1] * (kw - 1) + 1)) // stride[1] + 1
    return Ho, Wo


class _DCN(Function):
    """Shared forward/backward for the three variants."""

    @staticmethod
    def forward(ctx, x, offset, mask, weight, bias, stride, pad, dil, scales, groups, dg, out_from_offset):
        kh, kw = weight.shape[2:]
        if out_from_offset:   # pyramid: output grid is the offset grid (deform_conv.py:215-217)
            Ho, Wo = _out_hw(offset.shape[2], offset.shape[3], kh, kw, stride, pad, dil)
        else:
            Ho, Wo = _out_hw(x.shape[2], x.shape[3], kh, kw, stride, pad, dil)
        assert offset.shape[2] == Ho and offset.shape[3] == Wo, (offset.shape, Ho, Wo)
        ctx.cfg = (Ho, Wo, kh, kw, stride, pad, dil, scales, dg)
        ctx.groups = groups
        ctx.save_for_backward(x, offset, mask, weight, bias)
        B, Cout = x.shape[0], weight.shape[0]
        col = im2col(x, offset, mask, *ctx.cfg)                       # [C*kk, B, Ho, Wo]
        colg = col.view(groups, -1, B * Ho * Wo)
        wg = weight.reshape(groups, Cout // groups, -1)
        out = torch.bmm(wg, colg).view(Cout, B, Ho, Wo).permute(1, 0, 2, 3).contiguous()
        if bias is not None:
            out = out + bias.view(1, -1, 1, 1)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, offset, mask, weight, bias = ctx.saved_tensors
        Ho, Wo = ctx.cfg[:2]
        groups = ctx.groups
        B, Cout = gy.shape[:2]
        gyg = gy.permute(1, 0, 2, 3).reshape(groups, Cout // groups, B * Ho * Wo)
        wg = weight.reshape(groups, Cout // groups, -1)
        gcol = torch.bmm(wg.transpose(1, 2), gyg).reshape(-1, B, Ho, Wo)
        goff, gmask = col2im_coord(gcol, x, offset, mask, *ctx.cfg)
        gx = col2im(gcol, x, offset, mask, *ctx.cfg)
        col = im2col(x, offset, mask, *ctx.cfg).view(groups, -1, B * Ho * Wo)
        gw = torch.bmm(gyg, col.transpose(1, 2)).reshape(weight.shape)
        gb = gy.sum(dim=(0, 2, 3)) if bias is not None else None
        return gx, goff, gmask, gw, gb, None, None, None, None, None, None, None


def deform_conv(x, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1,
                im2col_step=64):
    """DCNv1 (deform_conv.py:15-111)."""
    return _DCN.apply(x, offset, None, weight, None, _pair(stride), _pair(padding), _pair(dilation),
                      (1.0, 1.0), groups, deformable_groups, False)


def modulated_deform_conv(x, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                          deformable_groups=1):
    """DCNv2 (deform_conv.py:114-185)."""
    return _DCN.apply(x, offset, mask, weight, bias, _pair(stride), _pair(padding), _pair(dilation),
                      (1.0, 1.0), groups, deformable_groups, False)


def pyramid_deform_conv(x, offset, weight, scales=1, stride=1, padding=0, dilation=1, groups=1,
                        deformable_groups=1, im2col_step=64):
    """LSNet pyramid DCN (deform_conv.py:188-287); scales = (scale_h, scale_w)."""
    scales = _pair(scales)
    return _DCN.apply(x, offset, None, weight, None, _pair(stride), _pair(padding), _pair(dilation),
                      (float(scales[0]), float(scales[1])), groups, deformable_groups, True)


# ---------------------------------------------------------------------------------------------
# Sigmoid focal loss (mmdet/ops/sigmoid_focal_loss/src/cuda/sigmoid_focal_loss_cuda.cu:23-97),
# restated with torch fp32 element-wise ops in the kernel's operation order.
# ---------------------------------------------------------------------------------------------
class _SigmoidFocal(Function):

    @staticmethod
    def forward(ctx, logits, targets, gamma, alpha):
        ctx.save_for_backward(logits, targets)
        ctx.gamma, ctx.alpha = gamma, alpha
        C = logits.shape[1]
        d = torch.arange(C, device=logits.device).view(1, C)
        t = targets.view(-1, 1)
        c1 = (t == d).to(logits.dtype)
        c2 = ((t >= 0) & (t != d)).to(logits.dtype)
        p = 1. / (1. + torch.exp(-logits))
        tiny = torch.finfo(torch.float32).tiny
        term1 = torch.pow(1. - p, gamma) * torch.log(p.clamp(min=tiny))
        ge = (logits >= 0).to(logits.dtype)
        term2 = torch.pow(p, gamma) * (-1. * logits * ge - torch.log(1. + torch.exp(logits - 2. * logits * ge)))
        return -c1 * term1 * alpha - c2 * term2 * (1.0 - alpha)

    @staticmethod
    @once_differentiable
    def backward(ctx, d_loss):
        logits, targets = ctx.saved_tensors
        gamma, alpha = ctx.gamma, ctx.alpha
        C = logits.shape[1]
        d = torch.arange(C, device=logits.device).view(1, C)
        t = targets.view(-1, 1)
        c1 = (t == d).to(logits.dtype)
        c2 = ((t >= 0) & (t != d)).to(logits.dtype)
        p = 1. / (1. + torch.exp(-logits))
        tiny = torch.finfo(torch.float32).tiny
        term1 = torch.pow(1. - p, gamma) * (1. - p - (p * gamma * torch.log(p.clamp(min=tiny))))
        ge = (logits >= 0).to(logits.dtype)
        term2 = torch.pow(p, gamma) * (
            (-1. * logits * ge - torch.log(1. + torch.exp(logits - 2. * logits * ge))) * (1. - p) * gamma - p)
        g = (-c1 * term1 * alpha - c2 * term2 * (1.0 - alpha)) * d_loss
        return g, None, None, None


def sigmoid_focal_loss_elementwise(logits, targets, gamma=2.0, alpha=0.25):
    return _SigmoidFocal.apply(logits, targets, float(gamma), float(alpha))
