"""Deformable convolutions on liblsnet_sm100.so with the reference's call signatures
(mmdet/ops/dcn/deform_conv.py:290-292): ``deform_conv``, ``modulated_deform_conv``, ``pyramid_deform_conv``.

Each autograd direction is ONE C-ABI call (include/lsnet_b200.h, "deformable convolution as whole operators"):
forward  : lsnet_dcn_forward — FUSED: bilinear-offset gather -> SWIZZLE_128B shared-memory A tile -> tcgen05.mma; no
           column matrix in HBM (shapes outside the fused kernel's range go gather -> GEMM through a workspace)
backward : lsnet_dcn_backward_data   dCol = dY x W (tcgen05) -> scatter to dX / reduce to dOffset, dMask
           lsnet_dcn_backward_weight dW += dY^T x columns (MN-major split-K tcgen05); the columns are either the
           forward's optional side output (LSNET_DCN_SAVE_COL=1: kept in HBM, not re-gathered as
           deform_conv_cuda.cpp:770-773 does) or re-sampled in the backward (LSNET_DCN_SAVE_COL=0: no column matrix
           survives the forward)
groups > 1 (the X-101 backbone sites: groups=64, widths 512/1024/2048): v1 runs the SAME kernels on the block-diagonal
expansion of the grouped weight (dense [Cout, taps*Cin] with zeros outside each group's channel block) and takes the
block diagonal of the dense weight gradient.  Exact: the extra products are all multiplications by zero.
"""
import torch
from torch.autograd import Function
from torch.nn.modules.utils import _pair

from .. import lib as L
from . import gemm_ops as G


def _pix_major(t, min_ld_mult=1):
    """(B,C,H,W) fp32 tensor -> (tensor, ld) with pixel-major memory (channel stride 1)."""
    B, C, H, W = t.shape
    ld = t.stride(3)
    ok = t.stride(1) == 1 and t.stride(2) == W * ld and (B == 1 or t.stride(0) == H * W * ld)
    if not ok or t.dtype != torch.float32:
        t = t.float().contiguous(memory_format=torch.channels_last)
        ld = t.stride(3)
    return t, ld


def dcn_im2col(x, offset, mask, Ho, Wo, kh, kw, stride, pad, dil, scales, dg, mask_logits=False):
    B, H, W, C, ldx = G.nhwc_geom(x)
    offset, ldo = _pix_major(offset)
    ldm = 0
    if mask is not None:
        mask, ldm = _pix_major(mask)
    col = torch.empty((B * Ho * Wo, kh * kw * C), device=x.device, dtype=torch.bfloat16)
    L.call('lsnet_dcn_im2col_bf16', L.ptr(x), L.c_int(B), L.c_int(H), L.c_int(W), L.c_int(C), L.c_ll(ldx),
           L.ptr(offset), L.c_ll(ldo), L.ptr(mask), L.c_ll(ldm), L.c_int(Ho), L.c_int(Wo), L.c_int(kh), L.c_int(kw),
           L.c_int(stride[0]), L.c_int(stride[1]), L.c_int(pad[0]), L.c_int(pad[1]), L.c_int(dil[0]), L.c_int(dil[1]),
           L.c_f(scales[0]), L.c_f(scales[1]), L.c_int(dg), L.ptr(col), L.c_ll(col.stride(0)), L.c_int(int(mask_logits)),
           L.stream())
    return col


import os

# accumulate dX in fp32 (v4.f32 reds) instead of packed bf16 (half the red instructions)
DX_FP32 = os.environ.get('LSNET_DCN_DX_FP32', '0') == '1'


def dcn_col2im(gcol, x, offset, mask, Ho, Wo, kh, kw, stride, pad, dil, scales, dg, need_dx=True, dx_fp32=None,
               mask_logits=False, packed_out=False):
    B, H, W, C, ldx = G.nhwc_geom(x)
    offset, ldo = _pix_major(offset)
    ldm = 0
    if mask is not None:
        mask, ldm = _pix_major(mask)
    taps = kh * kw
    dx_fp32 = DX_FP32 if dx_fp32 is None else dx_fp32
    dx = torch.zeros((B, H, W, C), device=x.device, dtype=torch.float32 if dx_fp32 else torch.bfloat16) if need_dx else None
    if packed_out:
        # ONE gradient tensor for the conv_offset output: dOffset in channels [0, 2*dg*taps), dMask (w.r.t. the
        # logits) behind it — the layout of ModulatedDeformConvPack's conv_offset (deform_conv.py:528-531)
        assert mask is not None
        nom = dg * 3 * taps
        dom = torch.empty((B, Ho, Wo, nom), device=x.device, dtype=torch.float32)
        doff, dmask, lddo, lddm = dom, dom[..., dg * 2 * taps:], nom, nom
    else:
        doff = torch.empty((B, Ho, Wo, dg * 2 * taps), device=x.device, dtype=torch.float32)
        dmask = torch.empty((B, Ho, Wo, dg * taps), device=x.device, dtype=torch.float32) if mask is not None else None
        lddo, lddm = dg * 2 * taps, dg * taps
    L.call('lsnet_dcn_col2im_bf16', L.ptr(gcol), L.c_ll(gcol.stride(0)), L.ptr(x), L.c_int(B), L.c_int(H), L.c_int(W),
           L.c_int(C), L.c_ll(ldx), L.ptr(offset), L.c_ll(ldo), L.ptr(mask), L.c_ll(ldm), L.c_int(Ho), L.c_int(Wo),
           L.c_int(kh), L.c_int(kw), L.c_int(stride[0]), L.c_int(stride[1]), L.c_int(pad[0]), L.c_int(pad[1]),
           L.c_int(dil[0]), L.c_int(dil[1]), L.c_f(scales[0]), L.c_f(scales[1]), L.c_int(dg), L.ptr(dx), L.c_ll(C),
           L.c_int(int(bool(dx_fp32))), L.ptr(doff), L.c_ll(lddo), L.ptr(dmask), L.c_ll(lddm),
           L.c_int(int(mask_logits)), L.stream())
    if packed_out:
        return dx.permute(0, 3, 1, 2) if dx is not None else None, dom.permute(0, 3, 1, 2), None
    return (dx.permute(0, 3, 1, 2) if dx is not None else None, doff.permute(0, 3, 1, 2),
            dmask.permute(0, 3, 1, 2) if dmask is not None else None)


# 1: the fused forward also writes the bf16 column matrix (side output) and the weight gradient reads it back;
# 0: nothing of the column matrix survives the forward, the weight gradient re-samples x.
SAVE_COL = os.environ.get('LSNET_DCN_SAVE_COL', '1') == '1'


def _desc(x_geom, Ho, Wo, kh, kw, stride, pad, dil, scales, dg, mask_logits, groups=1):
    B, H, W, C, ldx = x_geom
    return L.DcnDesc(B, H, W, C, ldx, Ho, Wo, kh, kw, stride[0], stride[1], pad[0], pad[1], dil[0], dil[1],
                     float(scales[0]), float(scales[1]), groups, dg, int(bool(mask_logits)), L.DTYPE_BF16)


def grouped_native(C, N, kh, kw, groups, dg):
    """True when the library runs this grouped weight on its block-diagonal kernels (lsnet_dcn_grouped_supported)."""
    import ctypes
    if groups <= 1:
        return False
    d = L.DcnDesc(1, 8, 8, C, C, 8, 8, kh, kw, 1, 1, 1, 1, 1, 1, 1.0, 1.0, groups, dg, 0, L.DTYPE_BF16)
    return bool(L.load().lsnet_dcn_grouped_supported(ctypes.byref(d), L.c_int(N)))


def pack_grouped_fwd(w, groups):
    """(C, C/groups, kh, kw) -> bf16 [C, kh*kw*64]: row o, column tap*64 + j = weight of out channel o for input channel
    (o // 64) * 64 + j (zero outside o's group)."""
    C, cpg, kh, kw = w.shape
    taps = kh * kw
    o = torch.arange(C, device=w.device)
    j0 = (o // cpg) * cpg - (o // 64) * 64
    out = torch.zeros((C, taps, 64), device=w.device, dtype=torch.bfloat16)
    jj = j0[:, None, None] + torch.arange(cpg, device=w.device)[None, None, :]
    out[o[:, None, None], torch.arange(taps, device=w.device)[None, :, None], jj] = \
        w.permute(0, 2, 3, 1).reshape(C, taps, cpg).to(torch.bfloat16)
    return out.view(C, taps * 64)


def pack_grouped_bwd(w, groups):
    """bf16 [kh*kw*C, 64]: row tap*C + blk*64 + j, column o' = weight of out channel blk*64 + o' for input channel
    blk*64 + j."""
    C, cpg, kh, kw = w.shape
    taps = kh * kw
    wg = pack_grouped_fwd(w, groups).view(C // 64, 64, taps, 64)          # [blk, o', tap, j]
    return wg.permute(2, 0, 3, 1).reshape(taps * C, 64).contiguous()


def unpack_grouped_dw(dwc, groups, kh, kw, dtype):
    """compact [C, kh*kw*256] fp32 (row o: its 256-channel tile per tap) -> (C, C/groups, kh, kw)."""
    C = dwc.shape[0]
    cpg, taps = C // groups, kh * kw
    o = torch.arange(C, device=dwc.device)
    j0 = ((o % 256) // cpg) * cpg
    jj = j0[:, None, None] + torch.arange(cpg, device=dwc.device)[None, None, :]
    d = dwc.view(C, taps, 256)[o[:, None, None], torch.arange(taps, device=dwc.device)[None, :, None], jj]   # (C,taps,cpg)
    return d.permute(0, 2, 1).reshape(C, cpg, kh, kw).to(dtype)


def _workspace(nbytes, device):
    return torch.empty(int(nbytes), device=device, dtype=torch.uint8) if nbytes else None


def dcn_forward(x, offset, mask, wp, bias, Ho, Wo, kh, kw, stride, pad, dil, scales, dg, mask_logits=False,
                out=None, out_dtype=torch.bfloat16, relu=False, save_col=False, groups=1, gn_sums=None, gn_groups=0):
    """Whole forward operator.  x (B,C,H,W) channels_last bf16; wp bf16 [Npad16, kh*kw*C]; returns (out2d [P, N], col or
    None).  ``out``: write into this [P, N] view (row pitch = out.stride(0))."""
    import ctypes
    geom = G.nhwc_geom(x)
    B, H, W, C, ldx = geom
    offset, ldo = _pix_major(offset)
    ldm = 0
    if mask is not None:
        mask, ldm = _pix_major(mask)
    N, P = wp.shape[0], B * Ho * Wo
    assert wp.dtype == torch.bfloat16 and wp.is_contiguous() and wp.shape[1] == kh * kw * (64 if groups > 1 else C)
    if out is None:
        out = torch.empty((P, N), device=x.device, dtype=out_dtype)
    assert out.stride(1) == 1 and out.dtype in (torch.bfloat16, torch.float32)
    d = _desc(geom, Ho, Wo, kh, kw, stride, pad, dil, scales, dg, mask_logits, groups)
    col = torch.empty((P, kh * kw * C), device=x.device, dtype=torch.bfloat16) if save_col else None
    ws_bytes = 0 if save_col else L.load().lsnet_dcn_forward_workspace_size(ctypes.byref(d), L.c_int(N))
    ws = _workspace(ws_bytes, x.device)
    # gn_sums: fp64 workspace of the GroupNorm that follows (zeroed); the epilogue adds the output's (sum, sum of squares)
    L.call('lsnet_dcn_forward_gn', ctypes.byref(d), L.ptr(x), L.ptr(offset), L.c_ll(ldo), L.ptr(mask), L.c_ll(ldm),
           L.ptr(wp), L.c_int(N), L.ptr(bias), L.c_int(int(relu)), L.ptr(out), L.c_ll(out.stride(0)),
           L.c_int(int(out.dtype == torch.float32)), L.ptr(col), L.ptr(ws), ctypes.c_size_t(ws_bytes), L.ptr(gn_sums),
           L.c_int(gn_groups), L.stream())
    return out, col


def dcn_backward_data(gy2, wt, x, offset, mask, Ho, Wo, kh, kw, stride, pad, dil, scales, dg, need_dx=True,
                      dx_fp32=None, mask_logits=False, packed_out=False, groups=1):
    """Whole backward-data operator: gy2 bf16 [P, N]; wt bf16 [kh*kw*C, N].  Returns (dx, doffset, dmask) as logical
    NCHW views (see dcn_col2im for ``packed_out``)."""
    import ctypes
    geom = G.nhwc_geom(x)
    B, H, W, C, ldx = geom
    offset, ldo = _pix_major(offset)
    ldm = 0
    if mask is not None:
        mask, ldm = _pix_major(mask)
    taps = kh * kw
    N = gy2.shape[1]
    assert gy2.dtype == torch.bfloat16 and gy2.stride(1) == 1 and wt.is_contiguous() and \
        wt.shape == (taps * C, 64 if groups > 1 else N)
    dx_fp32 = DX_FP32 if dx_fp32 is None else dx_fp32
    dx = torch.zeros((B, H, W, C), device=x.device, dtype=torch.float32 if dx_fp32 else torch.bfloat16) if need_dx else None
    if packed_out:
        assert mask is not None
        nom = dg * 3 * taps
        dom = torch.empty((B, Ho, Wo, nom), device=x.device, dtype=torch.float32)
        doff, dmask, lddo, lddm = dom, dom[..., dg * 2 * taps:], nom, nom
    else:
        doff = torch.empty((B, Ho, Wo, dg * 2 * taps), device=x.device, dtype=torch.float32)
        dmask = torch.empty((B, Ho, Wo, dg * taps), device=x.device, dtype=torch.float32) if mask is not None else None
        lddo, lddm = dg * 2 * taps, dg * taps
    d = _desc(geom, Ho, Wo, kh, kw, stride, pad, dil, scales, dg, mask_logits, groups)
    ws_bytes = L.load().lsnet_dcn_backward_data_workspace_size(ctypes.byref(d), L.c_int(N))
    ws = _workspace(ws_bytes, x.device)
    L.call('lsnet_dcn_backward_data', ctypes.byref(d), L.ptr(gy2), L.c_ll(gy2.stride(0)), L.c_int(N), L.ptr(wt), L.ptr(x),
           L.ptr(offset), L.c_ll(ldo), L.ptr(mask), L.c_ll(ldm), L.ptr(dx), L.c_ll(C), L.c_int(int(bool(dx_fp32))),
           L.ptr(doff), L.c_ll(lddo), L.ptr(dmask), L.c_ll(lddm), L.ptr(ws), ctypes.c_size_t(ws_bytes), L.stream())
    if packed_out:
        return dx.permute(0, 3, 1, 2) if dx is not None else None, dom.permute(0, 3, 1, 2), None
    return (dx.permute(0, 3, 1, 2) if dx is not None else None, doff.permute(0, 3, 1, 2),
            dmask.permute(0, 3, 1, 2) if dmask is not None else None)


def dcn_backward_weight(gy2, x, offset, mask, col, Ho, Wo, kh, kw, stride, pad, dil, scales, dg, mask_logits=False,
                        out=None, groups=1):
    """dW [N, kh*kw*C] fp32 (+= into ``out``) from gy2 bf16 [P, N] and either the saved columns or a re-sampling of x.
    groups > 1: the compact block-diagonal form [N, kh*kw*256] (unpack_grouped_dw)."""
    import ctypes
    geom = G.nhwc_geom(x)
    B, H, W, C, ldx = geom
    offset, ldo = _pix_major(offset)
    ldm = 0
    if mask is not None:
        mask, ldm = _pix_major(mask)
    N, K = gy2.shape[1], kh * kw * (256 if groups > 1 else C)
    if out is None:
        out = torch.zeros((N, K), device=x.device, dtype=torch.float32)
    assert out.dtype == torch.float32 and out.stride(1) == 1
    d = _desc(geom, Ho, Wo, kh, kw, stride, pad, dil, scales, dg, mask_logits, groups)
    ws_bytes = L.load().lsnet_dcn_backward_weight_workspace_size(ctypes.byref(d), L.c_int(N), L.c_int(int(col is not None)))
    ws = _workspace(ws_bytes, x.device)
    L.call('lsnet_dcn_backward_weight', ctypes.byref(d), L.ptr(gy2), L.c_ll(gy2.stride(0)), L.c_int(N), L.ptr(x),
           L.ptr(offset), L.c_ll(ldo), L.ptr(mask), L.c_ll(ldm), L.ptr(col), L.ptr(out), L.c_ll(out.stride(0)),
           L.ptr(ws), ctypes.c_size_t(ws_bytes), L.stream())
    return out


def _out_hw(H, W, kh, kw, stride, pad, dil):
    return ((H + 2 * pad[0] - (dil[0] * (kh - 1) + 1)) // stride[0] + 1,
            (W + 2 * pad[1] - (dil[1] * (kw - 1) + 1)) // stride[1] + 1)


OVERLAP_WGRAD = os.environ.get('LSNET_OVERLAP_WGRAD', '0') == '1'   # measured with the binned adjoint: 34.5 ms with, 32.4 ms without
_SIDE = {}


def _side_stream(device):
    key = str(device)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


def _expand_groups(w, groups):
    """(Cout, Cin/groups, kh, kw) grouped weight -> dense (Cout, Cin, kh, kw), zero outside each group's input block."""
    if groups == 1:
        return w
    co, cig, kh, kw = w.shape
    idx = torch.arange(groups, device=w.device)
    d = w.new_zeros(groups, co // groups, groups, cig, kh, kw)
    d[idx, :, idx] = w.reshape(groups, co // groups, cig, kh, kw)
    return d.view(co, groups * cig, kh, kw)


class _DCN(Function):

    @staticmethod
    def forward(ctx, x, offset, mask, weight, bias, stride, pad, dil, scales, groups, dg, out_from_offset, out_fp32,
                out_slice=None, packed_om=False, gx_sink=None, gn_holder=None, gn_groups=0, skip_bias_grad=False):
        ctx.skip_bias_grad = skip_bias_grad      # the GroupNorm behind this op adds the bias gradient (ops/norm.py bias_sink)
        ctx.gx_sink = gx_sink        # see ops/conv.py::_ConvSame: the conv_offset conv adds its input gradient into ours
        # gn_holder (dict) + gn_groups: a GroupNorm(gn_groups) follows; its statistics are accumulated by this op's GEMM
        # epilogue into holder['sums'] (ops/norm.py picks them up instead of running its own statistics pass)
        # packed_om: ``offset`` is the whole conv_offset output (B, 3*dg*taps, Ho, Wo) — offsets in the first 2/3 of the
        # channels, mask LOGITS behind them; ``mask`` is None and the kernels apply the sigmoid.
        co, cig, kh, kw = weight.shape
        ci = cig * groups
        if co % groups:
            raise ValueError(f'out_channels {co} is not divisible by groups {groups}')
        x = G.as_nhwc(x, torch.bfloat16)
        B, _, H, W = x.shape
        src = offset if out_from_offset else x
        Ho, Wo = _out_hw(src.shape[2], src.shape[3], kh, kw, stride, pad, dil)
        if offset.shape[2] != Ho or offset.shape[3] != Wo:
            raise ValueError(f'offset grid {tuple(offset.shape[2:])} != output grid {(Ho, Wo)}')
        cfg = (Ho, Wo, kh, kw, stride, pad, dil, scales, dg)
        ctx.packed_om = packed_om
        npad = (co + 15) // 16 * 16
        # grouped weights: the library's block-diagonal kernels when the shape qualifies (X-101 sites), else the dense
        # kernels on the zero-expanded weight
        ctx.native_groups = native = groups if grouped_native(ci, co, kh, kw, groups, dg) else 1

        def pack_fwd(t):
            if native > 1:
                return pack_grouped_fwd(t, groups)
            p = _expand_groups(t, groups).permute(0, 2, 3, 1).reshape(co, kh * kw * ci).to(torch.bfloat16)
            if npad != co:
                p = torch.cat([p, p.new_zeros(npad - co, p.shape[1])], 0)
            return p.contiguous()
        wp = G.cached_pack(weight, 'dcn_fwd%d_%d' % (groups, native), pack_fwd)
        b = None
        if bias is not None:
            b = G.cached_pack(bias, 'bias%d' % npad, lambda t: torch.cat([t.float(), t.new_zeros(npad - co).float()]))
        if packed_om:
            offset, _ = _pix_major(offset.detach())
            n2 = 2 * dg * kh * kw
            off_v, mask_v, logits = offset[:, :n2], offset[:, n2:], True
        else:
            off_v, mask_v, logits = offset.detach(), None if mask is None else mask.detach(), False
        save_col = SAVE_COL and weight.requires_grad and native == 1
        out2d = None
        if out_slice is not None:
            # write straight into channels [c0, c0+co) of a wider pixel-major buffer (replaces a later torch.cat)
            buf, c0 = out_slice
            assert npad == co and buf.dtype == torch.bfloat16 and buf.shape[:3] == (B, Ho, Wo) and c0 % 8 == 0
            ld = buf.shape[3]
            out2d = torch.as_strided(buf, (B * Ho * Wo, co), (ld, 1), buf.storage_offset() + c0)
        gn_sums = None
        if gn_holder is not None and gn_groups > 0 and native == 1 and not out_fp32 and out_slice is None and npad == co \
                and co % gn_groups == 0 and (co // gn_groups) % 8 == 0:
            gn_sums = torch.zeros(3 * B * gn_groups + 1, device=x.device, dtype=torch.float64)
            gn_holder['sums'] = gn_sums
        out2d, col = dcn_forward(x, off_v, mask_v, wp, b, *cfg, mask_logits=logits, out=out2d,
                                 out_dtype=torch.float32 if out_fp32 else torch.bfloat16, save_col=save_col, groups=native,
                                 gn_sums=gn_sums, gn_groups=gn_groups)
        ctx.save_for_backward(x, offset, mask, weight, col)
        ctx.cfg, ctx.has_bias, ctx.groups = cfg, bias is not None, groups
        ctx.grad2d = G.direct_grad(weight) if groups == 1 else None
        ctx.bias_param = bias
        if out_slice is not None:
            return torch.as_strided(buf, (B, co, Ho, Wo), (Ho * Wo * ld, 1, Wo * ld, ld), buf.storage_offset() + c0)
        return out2d.view(B, Ho, Wo, npad).permute(0, 3, 1, 2)[:, :co]

    @staticmethod
    def backward(ctx, gy):
        x, offset, mask, weight, col = ctx.saved_tensors
        Ho, Wo, kh, kw = ctx.cfg[:4]
        groups = ctx.groups
        co, ci = weight.shape[0], weight.shape[1] * groups
        B = x.shape[0]

        def unpack_dw(dw):       # dense [cop, taps*ci] fp32 -> (co, ci/groups, kh, kw): block diagonal when grouped
            d = dw[:co].view(co, kh, kw, ci)
            if groups > 1:
                idx = torch.arange(groups, device=d.device)
                d = d.view(groups, co // groups, kh, kw, groups, ci // groups)[idx, :, :, :, idx]
                d = d.reshape(co, kh, kw, ci // groups)
            return d.permute(0, 3, 1, 2).to(weight.dtype)
        gyp, colsum = G.grad_prep(gy, None, ctx.has_bias and ctx.needs_input_grad[4] and not ctx.skip_bias_grad,
                                  colsum_into=G.direct_vec(ctx.bias_param))
        cop = gyp.shape[1]
        gy2 = torch.as_strided(gyp, (B * Ho * Wo, cop), (gyp.stride(3), 1))
        gx = goff = gmask = gw = gb = None
        if ctx.packed_om:
            n2 = 2 * ctx.cfg[8] * kh * kw
            off_v, mask_v, logits = offset[:, :n2], offset[:, n2:], True
        else:
            off_v, mask_v, logits = offset.detach(), None if mask is None else mask.detach(), False

        native = ctx.native_groups

        def wgrad():
            if native > 1:
                dwc = dcn_backward_weight(gy2, x, off_v, mask_v, None, *ctx.cfg, mask_logits=logits, groups=native)
                return unpack_grouped_dw(dwc[:co], native, kh, kw, weight.dtype)
            direct = ctx.grad2d
            if direct is not None and tuple(direct.shape) != (cop, kh * kw * ci):
                direct = None
            if direct is not None:       # accumulate straight into the parameter's (tap-major) gradient memory
                dcn_backward_weight(gy2, x, off_v, mask_v, col, *ctx.cfg, mask_logits=logits, out=direct)
                return None
            return unpack_dw(dcn_backward_weight(gy2, x, off_v, mask_v, col, *ctx.cfg, mask_logits=logits))
        # The weight-gradient GEMM (tensor-pipe bound) only needs dY and the columns, the scatter (issue bound on the CUDA
        # cores) only needs dCol: optionally run the former on a side stream so the two overlap on the SMs.
        side = None
        if OVERLAP_WGRAD and ctx.needs_input_grad[3]:
            cur = torch.cuda.current_stream()
            side = _side_stream(x.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                gw = wgrad()
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or (mask is not None and ctx.needs_input_grad[2]):
            # B operand [N = taps*ci, K = co]: W^T, K-major
            def pack_bwd(t):
                if native > 1:
                    return pack_grouped_bwd(t, groups)
                p = _expand_groups(t, groups).permute(2, 3, 1, 0).reshape(kh * kw * ci, co).to(torch.bfloat16)
                if cop != co:
                    p = torch.cat([p, p.new_zeros(p.shape[0], cop - co)], 1)
                return p.contiguous()
            wt = G.cached_pack(weight, 'dcn_bwd%d_%d_%d' % (cop, groups, native), pack_bwd)
            gx, goff, gmask = dcn_backward_data(gy2, wt, x, off_v, mask_v, *ctx.cfg, need_dx=ctx.needs_input_grad[0],
                                                mask_logits=logits, packed_out=ctx.packed_om, groups=native)
            if gx is not None and gx.dtype != torch.bfloat16:
                gx = gx.to(torch.bfloat16)
            if gx is not None and ctx.gx_sink is not None:
                base = gx.permute(0, 2, 3, 1)
                if base.is_contiguous():
                    ctx.gx_sink['dx'] = base
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        elif ctx.needs_input_grad[3]:
            gw = wgrad()
        if ctx.has_bias and ctx.needs_input_grad[4] and not ctx.skip_bias_grad:
            gb = colsum
        return gx, goff, gmask, gw, gb, None, None, None, None, None, None, None, None, None, None, None, None, None, None


def deform_conv(x, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1, im2col_step=64,
                out_fp32=False):
    """DCNv1 — DeformConvFunction.apply (mmdet/ops/dcn/deform_conv.py:15-111)."""
    return _DCN.apply(x, offset, None, weight, None, _pair(stride), _pair(padding), _pair(dilation), (1.0, 1.0),
                      groups, deformable_groups, False, out_fp32)


def modulated_deform_conv(x, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                          deformable_groups=1, out_fp32=False):
    """DCNv2 — ModulatedDeformConvFunction.apply (mmdet/ops/dcn/deform_conv.py:114-185)."""
    return _DCN.apply(x, offset, mask, weight, bias, _pair(stride), _pair(padding), _pair(dilation), (1.0, 1.0),
                      groups, deformable_groups, False, out_fp32)


def modulated_deform_conv_packed(x, offset_mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                                 deformable_groups=1, out_fp32=False, gx_sink=None, gn_holder=None, gn_groups=0,
                                 skip_bias_grad=False):
    """ModulatedDeformConvPack.forward after its conv_offset (deform_conv.py:528-533) as ONE op: ``offset_mask`` is the
    raw conv_offset output; chunk / cat / sigmoid and their backward happen inside the sampling kernels."""
    return _DCN.apply(x, offset_mask, None, weight, bias, _pair(stride), _pair(padding), _pair(dilation), (1.0, 1.0),
                      groups, deformable_groups, False, out_fp32, None, True, gx_sink, gn_holder, gn_groups, skip_bias_grad)


def pyramid_deform_conv(x, offset, weight, scales=1, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1,
                        im2col_step=64, out_fp32=False, out_slice=None):
    """LSNet pyramid DCN — PyramidDeformConvFunction.apply (mmdet/ops/dcn/deform_conv.py:188-287); scales =
    (scale_h, scale_w); the output grid is the offset grid (:215-217)."""
    scales = _pair(scales)
    return _DCN.apply(x, offset, None, weight, None, _pair(stride), _pair(padding), _pair(dilation),
                      (float(scales[0]), float(scales[1])), groups, deformable_groups, True, out_fp32, out_slice)


class _JoinSlices(Function):
    """Autograd glue for outputs that several ops wrote into channel slices of ONE buffer: forward returns the buffer
    (no copy — this is what torch.cat would have produced), backward hands each producer its slice of the gradient."""

    @staticmethod
    def forward(ctx, buf, *slices):
        ctx.widths = [s.shape[1] for s in slices]
        return buf.view_as(buf)

    @staticmethod
    def backward(ctx, g):
        outs, c = [], 0
        for w in ctx.widths:
            outs.append(g[:, c:c + w])
            c += w
        return (None,) + tuple(outs)


def join_slices(buf_nchw, slices):
    return _JoinSlices.apply(buf_nchw, *slices)
