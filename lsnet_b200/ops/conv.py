"""Autograd convolution on the tcgen05 implicit-GEMM kernels: stride-1 'same' k x k convs and 1x1 convs over
pixel-major (channels_last) bf16 activations.  Forward, input-gradient (same kernel, flipped/transposed weights) and
weight-gradient (MN-major split-K kernel) all run in liblsnet_sm100.so; only the tiny bias-gradient column sum and the
weight re-packing use torch ops."""
import torch
from torch.autograd import Function

from . import gemm_ops as G


class _ConvSame(Function):

    @staticmethod
    def forward(ctx, x, weight, bias, pad, dil, relu, out_fp32):
        co, ci, kh, kw = weight.shape
        x = G.as_nhwc(x, torch.bfloat16)
        if x.shape[1] % 8:
            x = G.pad_channels_nhwc(x, 8)
        wp = G.cached_pack(weight, 'fwd', G.pack_conv_weight)
        npad = wp.shape[0]
        b = None
        if bias is not None:
            b = G.cached_pack(bias, 'bias%d' % npad, lambda t: torch.cat([t.float(), t.new_zeros(npad - co).float()]))
        out = G.conv2d_nhwc(x, wp, kh, kw, pad, dil, b, relu, torch.float32 if out_fp32 else torch.bfloat16,
                            n_valid=co)
        ctx.save_for_backward(x, weight, out if (relu and not out_fp32) else None)
        assert not (relu and out_fp32)
        ctx.cfg = (pad, dil, relu, bias is not None, ci)
        ctx.grad2d = G.direct_grad(weight)
        ctx.bias_param = bias
        return out

    @staticmethod
    def backward(ctx, gy):
        x, weight, out = ctx.saved_tensors
        pad, dil, relu, has_bias, ci = ctx.cfg
        co, _, kh, kw = weight.shape
        # one pass: cast to bf16, pad channels to 8, apply the ReLU mask, column-sum for the bias gradient
        # the bias gradient is added straight into the parameter's gradient memory when the trainer exposes it
        gyp, colsum = G.grad_prep(gy, out if relu else None, has_bias and ctx.needs_input_grad[2],
                                  colsum_into=G.direct_vec(ctx.bias_param))
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            wt = G.cached_pack(weight, 'bwd', lambda t: G.pack_conv_weight(t, flip_transpose=True))
            gx = G.conv2d_nhwc(gyp, wt, kh, kw, pad, dil, None, False, torch.bfloat16, n_valid=ci)
        if ctx.needs_input_grad[1]:
            direct = ctx.grad2d
            if direct is not None and tuple(direct.shape) != (gyp.shape[1], kh * kw * x.shape[1]):
                direct = None            # padded channel counts: go through the staging buffer
            if direct is not None:       # accumulate straight into the parameter's (tap-major) gradient memory
                G.conv2d_wgrad_nhwc(gyp, x, kh, kw, pad, dil, out=direct)
            else:
                dw = G.conv2d_wgrad_nhwc(gyp, x, kh, kw, pad, dil)       # (co_pad, taps, C_pad8)
                gw = dw[:co, :, :ci].reshape(co, kh, kw, ci).permute(0, 3, 1, 2).to(weight.dtype)
        if has_bias and ctx.needs_input_grad[2]:
            gb = colsum
        return gx, gw, gb, None, None, None, None


def conv2d_same(x, weight, bias=None, padding=1, dilation=1, relu=False, out_fp32=False):
    """nn.Conv2d(stride=1, padding=dilation*(k-1)/2) semantics on (B,C,H,W) tensors (any layout; converted to
    channels_last bf16).  Returns a (B,Cout,H,W) channels_last view."""
    return _ConvSame.apply(x, weight, bias, padding, dilation, relu, out_fp32)


def linear_nhwc(x, weight, bias=None, relu=False, out_fp32=False):
    """1x1 convolution."""
    return _ConvSame.apply(x, weight, bias, 0, 1, relu, out_fp32)
