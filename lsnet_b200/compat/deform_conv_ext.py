"""Drop-in for the reference's compiled ``deform_conv_ext`` module: the same eight functions with the pybind signatures of
mmdet/ops/dcn/src/deform_conv_ext.cpp:74-250 (argument order and in-place output convention included), implemented on
the whole-operator C ABI of liblsnet_sm100.so (include/lsnet_b200.h: lsnet_dcn_forward / _backward_data /
_backward_weight).  Replacing ``from . import deform_conv_ext`` in mmdet/ops/dcn/deform_conv.py:12 by
``from lsnet_b200.compat import deform_conv_ext`` puts the reference's autograd Functions (DeformConvFunction,
ModulatedDeformConvFunction, PyramidDeformConvFunction) on the B200 kernels unchanged.

The reference hands over NCHW fp32 tensors and preallocated outputs; the shim converts to the library's pixel-major bf16
layout at this edge and copies results back (the native modules in lsnet_b200/modules keep everything pixel-major and
skip these conversions).  ``columns`` / ``ones`` / ``im2col_step`` are scratch arguments of the reference's algorithm
and are ignored.  tests/test_gpu_compat_ext.py calls this module and the reference's own compiled extension with
identical arguments and compares every output."""
import ctypes

import torch

from .. import lib as L
from ..ops.dcn import _expand_groups


def _nhwc(t, dtype):
    return t.detach().to(dtype).permute(0, 2, 3, 1).contiguous()


def _desc(x, Ho, Wo, kH, kW, dH, dW, padH, padW, dilH, dilW, dg, scaleH=1.0, scaleW=1.0):
    B, H, W, C = x.shape
    return L.DcnDesc(B, H, W, C, C, Ho, Wo, kH, kW, dH, dW, padH, padW, dilH, dilW, float(scaleH), float(scaleW), 1, dg, 0,
                     L.DTYPE_BF16)


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 1), device=device, dtype=torch.uint8)


def _dense(weight, group):
    return _expand_groups(weight.detach(), group) if group > 1 else weight.detach()


def _forward(input, weight, bias, offset, mask, output, kH, kW, dH, dW, padH, padW, dilH, dilW, group, dg, scaleH=1.0,
             scaleW=1.0):
    if not input.is_cuda:
        raise NotImplementedError            # as the reference (deform_conv.py:46-47)
    w = _dense(weight, group)
    Co = w.shape[0]
    Np = (Co + 15) // 16 * 16
    Ho, Wo = output.shape[2:]
    x, off = _nhwc(input, torch.bfloat16), _nhwc(offset, torch.float32)
    msk = None if mask is None else _nhwc(mask, torch.float32)
    d = _desc(x, Ho, Wo, kH, kW, dH, dW, padH, padW, dilH, dilW, dg, scaleH, scaleW)
    wp = torch.zeros((Np, kH * kW * x.shape[3]), device=x.device, dtype=torch.bfloat16)
    wp[:Co] = w.permute(0, 2, 3, 1).reshape(Co, -1)
    b = None
    if bias is not None:
        b = torch.zeros(Np, device=x.device, dtype=torch.float32)
        b[:Co] = bias.detach().float()
    out = torch.empty((x.shape[0] * Ho * Wo, Np), device=x.device, dtype=torch.float32)
    n = L.load().lsnet_dcn_forward_workspace_size(ctypes.byref(d), L.c_int(Np))
    ws = _ws(n, x.device)
    L.call('lsnet_dcn_forward', ctypes.byref(d), L.ptr(x), L.ptr(off), L.c_ll(off.shape[-1]), L.ptr(msk),
           L.c_ll(0 if msk is None else msk.shape[-1]), L.ptr(wp), L.c_int(Np), L.ptr(b), L.c_int(0), L.ptr(out), L.c_ll(Np),
           L.c_int(1), L.ptr(None), L.ptr(ws), ctypes.c_size_t(n), L.stream())
    output.copy_(out.view(x.shape[0], Ho, Wo, Np)[..., :Co].permute(0, 3, 1, 2))


def _backward_data(input, weight, offset, mask, grad_output, grad_input, grad_offset, grad_mask, kH, kW, dH, dW, padH, padW,
                   dilH, dilW, group, dg, scaleH=1.0, scaleW=1.0):
    w = _dense(weight, group)
    Co = w.shape[0]
    Np = (Co + 7) // 8 * 8
    Ho, Wo = grad_output.shape[2:]
    x, off = _nhwc(input, torch.bfloat16), _nhwc(offset, torch.float32)
    msk = None if mask is None else _nhwc(mask, torch.float32)
    gy = torch.zeros((x.shape[0] * Ho * Wo, Np), device=x.device, dtype=torch.bfloat16)
    gy[:, :Co] = grad_output.detach().permute(0, 2, 3, 1).reshape(-1, Co)
    d = _desc(x, Ho, Wo, kH, kW, dH, dW, padH, padW, dilH, dilW, dg, scaleH, scaleW)
    wt = torch.zeros((kH * kW * x.shape[3], Np), device=x.device, dtype=torch.bfloat16)
    wt[:, :Co] = w.permute(2, 3, 1, 0).reshape(-1, Co)
    dx = torch.zeros(x.shape, device=x.device, dtype=torch.float32)          # accumulated by the kernel
    doff = torch.empty_like(off)
    dmsk = None if msk is None else torch.empty_like(msk)
    n = L.load().lsnet_dcn_backward_data_workspace_size(ctypes.byref(d), L.c_int(Np))
    ws = _ws(n, x.device)
    L.call('lsnet_dcn_backward_data', ctypes.byref(d), L.ptr(gy), L.c_ll(Np), L.c_int(Np), L.ptr(wt), L.ptr(x), L.ptr(off),
           L.c_ll(off.shape[-1]), L.ptr(msk), L.c_ll(0 if msk is None else msk.shape[-1]), L.ptr(dx), L.c_ll(x.shape[3]),
           L.c_int(1), L.ptr(doff), L.c_ll(off.shape[-1]), L.ptr(dmsk), L.c_ll(0 if msk is None else msk.shape[-1]),
           L.ptr(ws), ctypes.c_size_t(n), L.stream())
    grad_input.copy_(dx.permute(0, 3, 1, 2))
    grad_offset.copy_(doff.permute(0, 3, 1, 2))
    if grad_mask is not None and dmsk is not None:
        grad_mask.copy_(dmsk.permute(0, 3, 1, 2))


def _backward_weight(input, weight_shape, offset, mask, grad_output, grad_weight, kH, kW, dH, dW, padH, padW, dilH, dilW,
                     group, dg, scale=1.0, scaleH=1.0, scaleW=1.0):
    Co = weight_shape[0]
    Np = (Co + 7) // 8 * 8
    Ho, Wo = grad_output.shape[2:]
    x, off = _nhwc(input, torch.bfloat16), _nhwc(offset, torch.float32)
    msk = None if mask is None else _nhwc(mask, torch.float32)
    C = x.shape[3]
    gy = torch.zeros((x.shape[0] * Ho * Wo, Np), device=x.device, dtype=torch.bfloat16)
    gy[:, :Co] = grad_output.detach().permute(0, 2, 3, 1).reshape(-1, Co)
    d = _desc(x, Ho, Wo, kH, kW, dH, dW, padH, padW, dilH, dilW, dg, scaleH, scaleW)
    dw = torch.zeros((Np, kH * kW * C), device=x.device, dtype=torch.float32)
    n = L.load().lsnet_dcn_backward_weight_workspace_size(ctypes.byref(d), L.c_int(Np), L.c_int(0))
    ws = _ws(n, x.device)
    L.call('lsnet_dcn_backward_weight', ctypes.byref(d), L.ptr(gy), L.c_ll(Np), L.c_int(Np), L.ptr(x), L.ptr(off),
           L.c_ll(off.shape[-1]), L.ptr(msk), L.c_ll(0 if msk is None else msk.shape[-1]), L.ptr(None), L.ptr(dw),
           L.c_ll(dw.shape[1]), L.ptr(ws), ctypes.c_size_t(n), L.stream())
    dense = dw[:Co].view(Co, kH, kW, C).permute(0, 3, 1, 2)                 # (Co, C, kH, kW)
    if group > 1:                                                           # block diagonal of the dense gradient
        idx = torch.arange(group, device=dense.device)
        dense = dense.reshape(group, Co // group, group, C // group, kH, kW)[idx, :, idx].reshape(weight_shape)
    grad_weight.add_(scale * dense.to(grad_weight.dtype))                   # the reference accumulates (addmm_ beta = 1)


# ---- the eight entry points (deform_conv_ext.cpp:227-250) -----------------------------------------------------------
def deform_conv_forward(input, weight, offset, output, columns, ones, kW, kH, dW, dH, padW, padH, dilationW, dilationH,
                        group, deformable_group, im2col_step):
    _forward(input, weight, None, offset, None, output, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group,
             deformable_group)
    return 1


def deform_conv_backward_input(input, offset, gradOutput, gradInput, gradOffset, weight, columns, kW, kH, dW, dH, padW, padH,
                               dilationW, dilationH, group, deformable_group, im2col_step):
    _backward_data(input, weight, offset, None, gradOutput, gradInput, gradOffset, None, kH, kW, dH, dW, padH, padW,
                   dilationH, dilationW, group, deformable_group)
    return 1


def deform_conv_backward_parameters(input, offset, gradOutput, gradWeight, columns, ones, kW, kH, dW, dH, padW, padH,
                                    dilationW, dilationH, group, deformable_group, scale, im2col_step):
    _backward_weight(input, gradWeight.shape, offset, None, gradOutput, gradWeight, kH, kW, dH, dW, padH, padW, dilationH,
                     dilationW, group, deformable_group, scale)
    return 1


def modulated_deform_conv_forward(input, weight, bias, ones, offset, mask, output, columns, kernel_h, kernel_w, stride_h,
                                  stride_w, pad_h, pad_w, dilation_h, dilation_w, group, deformable_group, with_bias):
    _forward(input, weight, bias if with_bias else None, offset, mask, output, kernel_h, kernel_w, stride_h, stride_w, pad_h,
             pad_w, dilation_h, dilation_w, group, deformable_group)


def modulated_deform_conv_backward(input, weight, bias, ones, offset, mask, columns, grad_input, grad_weight, grad_bias,
                                   grad_offset, grad_mask, grad_output, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                                   dilation_h, dilation_w, group, deformable_group, with_bias):
    _backward_data(input, weight, offset, mask, grad_output, grad_input, grad_offset, grad_mask, kernel_h, kernel_w, stride_h,
                   stride_w, pad_h, pad_w, dilation_h, dilation_w, group, deformable_group)
    _backward_weight(input, weight.shape, offset, mask, grad_output, grad_weight, kernel_h, kernel_w, stride_h, stride_w,
                     pad_h, pad_w, dilation_h, dilation_w, group, deformable_group)
    if with_bias:
        grad_bias.add_(grad_output.detach().sum((0, 2, 3)).to(grad_bias.dtype))


def pyramid_deform_conv_forward(input, weight, offset, output, columns, ones, kW, kH, dW, dH, padW, padH, dilationW,
                                dilationH, scaleW, scaleH, group, deformable_group, im2col_step):
    _forward(input, weight, None, offset, None, output, kH, kW, dH, dW, padH, padW, dilationH, dilationW, group,
             deformable_group, scaleH, scaleW)
    return 1


def pyramid_deform_conv_backward_input(input, offset, gradOutput, gradInput, gradOffset, weight, columns, kW, kH, dW, dH,
                                       padW, padH, dilationW, dilationH, scaleW, scaleH, group, deformable_group,
                                       im2col_step):
    _backward_data(input, weight, offset, None, gradOutput, gradInput, gradOffset, None, kH, kW, dH, dW, padH, padW,
                   dilationH, dilationW, group, deformable_group, scaleH, scaleW)
    return 1


def pyramid_deform_conv_backward_parameters(input, offset, gradOutput, gradWeight, columns, ones, kW, kH, dW, dH, padW, padH,
                                            dilationW, dilationH, scaleW, scaleH, group, deformable_group, scale,
                                            im2col_step):
    _backward_weight(input, gradWeight.shape, offset, None, gradOutput, gradWeight, kH, kW, dH, dW, padH, padW, dilationH,
                     dilationW, group, deformable_group, scale, scaleH, scaleW)
    return 1
