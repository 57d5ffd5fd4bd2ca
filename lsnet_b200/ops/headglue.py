"""LSHead element-wise glue on liblsnet_sm100.so (lsnet_pred_reg_fwd/bwd, lsnet_add_softplus): softplus + get_pred_reg +
gradient-mul mix + base-offset subtraction as ONE kernel each way, and refine = softplus(raw + init.detach())
(mmdet/models/dense_heads/lsnet_head.py:372-400, 585-587, 735-755)."""
import math

import torch
from torch.autograd import Function

from .. import lib as L


def pred_reg_table(task_branch, num_vectors, num_kernel_points, n_out):
    """(n_sp, src[18], mode[18]) for LSHead.get_pred_reg.  'bbox': 10 signed slot pairs of the 20 softplus channels + the 8
    free offset channels; other branches: 8 selected contour / keypoint vectors + the centre, two signed pairs each."""
    n_off = 2 * num_kernel_points
    if task_branch == 'bbox':
        n_sp = 4 * (4 + 1)
        npair = n_sp // 2
        src = [2 * j for j in range(npair)] + [n_sp + k for k in range(n_off - npair)]
        mode = [0] * npair + [1] * (n_off - npair)
        assert n_out == n_sp + (n_off - npair)
    else:
        n_sp = n_out
        npts = n_out // 4                      # num_vectors + 1 (the last one is the centre)
        polys = list(range(npts - 1))
        sel = polys[::math.ceil(num_vectors / (num_kernel_points - 1))] if task_branch == 'segm' else polys[1::2]
        pts = sel + [npts - 1]
        assert 2 * len(pts) == n_off, (task_branch, len(pts), n_off)
        src = [4 * pts[j // 2] + 2 * (j % 2) for j in range(n_off)]
        mode = [0] * n_off
    return n_sp, src, mode


def _rows(t):
    """(B,C,H,W) fp32 pixel-major tensor -> (tensor, pixel pitch); copies only when the layout is something else."""
    B, C, H, W = t.shape
    ld = t.stride(3)
    if t.dtype != torch.float32 or t.stride(1) != 1 or t.stride(2) != W * ld or (B > 1 and t.stride(0) != H * W * ld):
        t = t.float().contiguous(memory_format=torch.channels_last)
        ld = t.stride(3)
    return t, ld


class _PredReg(Function):

    @staticmethod
    def forward(ctx, o, n_sp, src, mode, base, gradient_mul):
        o, ldo = _rows(o.detach())
        B, n_out, H, W = o.shape
        n_off = len(src)
        ldsp = (n_sp + 3) // 4 * 4
        sp = torch.empty((B, H, W, ldsp), device=o.device, dtype=torch.float32)
        off = torch.empty((B, H, W, n_off), device=o.device, dtype=torch.float32)
        tabs = (L.host_int_array(src), L.host_int_array(mode), L.host_float_array(base))
        L.call('lsnet_pred_reg_fwd', L.ptr(o), L.c_ll(ldo), L.c_ll(B * H * W), L.c_int(n_sp), L.c_int(n_out),
               L.c_int(n_off), *tabs, L.ptr(sp), L.c_ll(ldsp), L.ptr(off), L.c_ll(n_off), L.stream())
        ctx.save_for_backward(o)
        ctx.cfg = (n_sp, n_off, tabs, float(gradient_mul), ldo)
        return sp.permute(0, 3, 1, 2)[:, :n_sp], off.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gsp, goff):
        o, = ctx.saved_tensors
        n_sp, n_off, tabs, gm, ldo = ctx.cfg
        B, n_out, H, W = o.shape
        ldgsp = ldgoff = 0
        if gsp is not None:
            gsp, ldgsp = _rows(gsp)
        if goff is not None:
            goff, ldgoff = _rows(goff)
        go = torch.empty((B, H, W, n_out), device=o.device, dtype=torch.float32)
        L.call('lsnet_pred_reg_bwd', L.ptr(o), L.c_ll(ldo), L.c_ll(B * H * W), L.c_int(n_sp), L.c_int(n_out),
               L.c_int(n_off), *tabs, L.ptr(gsp), L.c_ll(ldgsp), L.ptr(goff), L.c_ll(ldgoff), L.c_f(gm), L.ptr(go),
               L.c_ll(n_out), L.stream())
        return go.permute(0, 3, 1, 2), None, None, None, None, None


def pred_reg(o, n_sp, src, mode, base, gradient_mul):
    """o (B, n_out, H, W) -> (softplus(o[:, :n_sp]), dcn offsets (B, len(src), H, W))."""
    return _PredReg.apply(o, n_sp, src, mode, base, gradient_mul)


class _AddSoftplus(Function):

    @staticmethod
    def forward(ctx, t, s):
        t, ldt = _rows(t.detach())
        s, lds = _rows(s.detach())
        B, C, H, W = t.shape
        ldo = (C + 3) // 4 * 4
        out = torch.empty((B, H, W, ldo), device=t.device, dtype=torch.float32)
        L.call('lsnet_add_softplus', L.ptr(t), L.c_ll(ldt), L.ptr(s), L.c_ll(lds), L.ptr(None), L.c_ll(0),
               L.c_ll(B * H * W), L.c_int(C), L.ptr(out), L.c_ll(ldo), L.stream())
        ctx.save_for_backward(t, s)
        return out.permute(0, 3, 1, 2)[:, :C]

    @staticmethod
    def backward(ctx, gy):
        t, s = ctx.saved_tensors
        B, C, H, W = t.shape
        gy, ldgy = _rows(gy)
        gt = torch.empty((B, H, W, C), device=t.device, dtype=torch.float32)
        L.call('lsnet_add_softplus', L.ptr(t), L.c_ll(t.stride(3)), L.ptr(s), L.c_ll(s.stride(3)), L.ptr(gy), L.c_ll(ldgy),
               L.c_ll(B * H * W), L.c_int(C), L.ptr(gt), L.c_ll(C), L.stream())
        return gt.permute(0, 3, 1, 2), None


def add_softplus(t, s_detached):
    """softplus(t + s) with no gradient into s (refine = softplus(raw + init.detach()))."""
    return _AddSoftplus.apply(t, s_detached)


class _UpsampleAdd(torch.autograd.Function):
    """fine + F.interpolate(coarse, size=fine.shape[2:], mode='nearest') in ONE kernel each way (FPN top-down pathway,
    mmdet/models/necks/fpn.py:180-192): lsnet_upsample_add_nhwc_bf16 / _bwd."""

    @staticmethod
    def forward(ctx, fine, coarse):
        from . import gemm_ops as G
        fine, coarse = G.as_nhwc(fine, torch.bfloat16), G.as_nhwc(coarse, torch.bfloat16)
        B, Hf, Wf, C, ldf = G.nhwc_geom(fine)
        _, Hc, Wc, Cc, ldc = G.nhwc_geom(coarse)
        assert C == Cc and C % 8 == 0
        out = torch.empty((B, Hf, Wf, C), device=fine.device, dtype=torch.bfloat16)
        L.call('lsnet_upsample_add_nhwc_bf16', L.ptr(fine), L.c_ll(ldf), L.ptr(coarse), L.c_ll(ldc), L.c_int(B), L.c_int(Hf),
               L.c_int(Wf), L.c_int(Hc), L.c_int(Wc), L.c_int(C), L.ptr(out), L.c_ll(C), L.stream())
        ctx.geom = (B, Hf, Wf, Hc, Wc, C)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        from . import gemm_ops as G
        B, Hf, Wf, Hc, Wc, C = ctx.geom
        g = G.as_nhwc(g, torch.bfloat16)
        gc = None
        if ctx.needs_input_grad[1]:
            gc = torch.empty((B, Hc, Wc, C), device=g.device, dtype=torch.bfloat16)
            L.call('lsnet_upsample_add_bwd_nhwc_bf16', L.ptr(g), L.c_ll(G.nhwc_geom(g)[4]), L.c_int(B), L.c_int(Hf), L.c_int(Wf),
                   L.c_int(Hc), L.c_int(Wc), L.c_int(C), L.ptr(gc), L.c_ll(C), L.stream())
            gc = gc.permute(0, 3, 1, 2)
        return (g if ctx.needs_input_grad[0] else None), gc


def upsample_add(fine, coarse):
    return _UpsampleAdd.apply(fine, coarse)
