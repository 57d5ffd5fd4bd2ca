"""Autograd convolution on the tcgen05 implicit-GEMM kernels: stride-1 'same' k x k convs and 1x1 convs over
pixel-major (channels_last) bf16 activations.  Forward, input-gradient (same kernel, flipped/transposed weights) and
weight-gradient (MN-major split-K kernel) all run in liblsnet_sm100.so; only the tiny bias-gradient column sum and the
weight re-packing use torch ops."""
import torch
from torch.autograd import Function

from .. import lib as L
from . import gemm_ops as G


class _ConvSame(Function):

    @staticmethod
    def forward(ctx, x, weight, bias, pad, dil, relu, out_fp32, gx_sink=None):
        # gx_sink: dict shared with ANOTHER consumer of x on the same stream whose backward runs first and leaves its input
        # gradient there under 'dx' ([B,H,W,C] bf16): this op then adds its own input gradient into that tensor in the
        # GEMM epilogue and returns none (the DCN + conv_offset pair: no separate gradient-sum kernel)
        ctx.gx_sink = gx_sink
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
            acc = ctx.gx_sink.pop('dx', None) if ctx.gx_sink is not None else None
            B_, _, H_, W_ = gyp.shape
            if acc is not None and acc.shape == (B_, H_, W_, wt.shape[0]) and acc.dtype == torch.bfloat16 \
                    and acc.is_contiguous() and wt.shape[0] == ci:
                taps = [(ky * dil - pad, kx * dil - pad, ky * kw + kx) for ky in range(kh) for kx in range(kw)]
                conv2d_taps(gyp, wt, taps, H_, W_, out=acc, resid=acc.permute(0, 3, 1, 2))     # acc += dgrad, in place
            else:
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
        return gx, gw, gb, None, None, None, None, None


def conv2d_same(x, weight, bias=None, padding=1, dilation=1, relu=False, out_fp32=False, gx_sink=None):
    """nn.Conv2d(stride=1, padding=dilation*(k-1)/2) semantics on (B,C,H,W) tensors (any layout; converted to
    channels_last bf16).  Returns a (B,Cout,H,W) channels_last view."""
    return _ConvSame.apply(x, weight, bias, padding, dilation, relu, out_fp32, gx_sink)


def linear_nhwc(x, weight, bias=None, relu=False, out_fp32=False):
    """1x1 convolution."""
    return _ConvSame.apply(x, weight, bias, 0, 1, relu, out_fp32)


# ---------------------------------------------------------------------------------------------------------------
# General (strided) convolution with a pre-packed bf16 weight: the trunk's Bottleneck convs with the frozen BatchNorm
# folded in (mmdet/models/backbones/resnet.py:261-301) and FPN's stride-2 extra levels (necks/fpn.py:203-211).
# ---------------------------------------------------------------------------------------------------------------
def conv_out_hw(H, W, kh, kw, stride, pad, dil):
    return ((H + 2 * pad - dil * (kh - 1) - 1) // stride + 1, (W + 2 * pad - dil * (kw - 1) - 1) // stride + 1)


def conv2d_taps(x, wp, taps, Ho, Wo, ish=1, isw=1, wt_taps=None, out=None, out_map=None, bias=None, resid=None,
                mask=None, relu=False, out_dtype=torch.bfloat16):
    """lsnet_conv2d_taps_bf16.  x (B,C,H,W) channels_last bf16; wp bf16 [N, wt_taps*Cpad64]; taps = [(dy, dx, kblk)];
    out_map = (osh, osw, ooh, oow, OH, OW) (default: the Ho x Wo grid itself).  Returns the logical (B, N, OH, OW) view."""
    B, H, W, C, ldp = G.nhwc_geom(x)
    N = wp.shape[0]
    osh, osw, ooh, oow, OH, OW = out_map or (1, 1, 0, 0, Ho, Wo)
    if out is None:
        out = torch.empty((B, OH, OW, N), device=x.device, dtype=out_dtype)
    assert out.shape == (B, OH, OW, N) and out.is_contiguous()
    wt_taps = wt_taps or len(taps)
    assert wp.dtype == torch.bfloat16 and wp.is_contiguous() and wp.shape[1] == wt_taps * ((C + 63) // 64 * 64)
    ldr = ldm = 0
    if resid is not None:
        rb, rh, rw, rc, ldr = G.nhwc_geom(resid)
        assert (rb, rh, rw) == (B, OH, OW) and rc >= N and resid.dtype == torch.bfloat16
    if mask is not None:
        mb, mh, mw, mc, ldm = G.nhwc_geom(mask)
        assert (mb, mh, mw) == (B, OH, OW) and mc >= N and mask.dtype == torch.bfloat16
    if len(taps) == 1 and tuple(taps[0]) == (0, 0, 0) and wt_taps == 1 and ish == isw == 1 and C % 64 == 0 \
            and (osh, osw, ooh, oow, OH, OW) == (1, 1, 0, 0, H, W) and (Ho, Wo) == (H, W):
        # 1x1 / stride 1: the plain [pixels, C] GEMM (no spatial patches, so no tile padding on ragged maps)
        L.call('lsnet_gemm_ex_bf16', L.ptr(x), L.c_ll(ldp), L.ptr(wp), L.c_ll(wp.stride(0)), L.ptr(out), L.c_ll(N),
               L.c_int(B * H * W), L.c_int(N), L.c_int(C), L.ptr(bias), L.ptr(resid), L.c_ll(ldr), L.ptr(mask), L.c_ll(ldm),
               L.c_int(int(relu)), L.c_int(int(out.dtype == torch.float32)), L.stream())
        return out.permute(0, 3, 1, 2)
    L.call('lsnet_conv2d_taps_bf16', L.ptr(x), L.c_int(B), L.c_int(H), L.c_int(W), L.c_int(C), L.c_ll(ldp), L.ptr(wp),
           L.c_int(N), L.c_int(len(taps)), L.host_int_array([t[0] for t in taps]), L.host_int_array([t[1] for t in taps]),
           L.host_int_array([t[2] for t in taps]), L.c_int(wt_taps), L.c_int(ish), L.c_int(isw), L.c_int(Ho), L.c_int(Wo),
           L.ptr(out), L.c_ll(N), L.c_int(osh), L.c_int(osw), L.c_int(ooh), L.c_int(oow), L.c_int(OH), L.c_int(OW),
           L.ptr(bias), L.ptr(resid), L.c_ll(ldr), L.ptr(mask), L.c_ll(ldm), L.c_int(int(relu)),
           L.c_int(int(out.dtype == torch.float32)), L.stream())
    return out.permute(0, 3, 1, 2)


def dgrad_phases(kh, kw, stride, pad, dil):
    """Input gradient of a strided convolution as one implicit GEMM per output phase: phase (ph, pw) holds the input pixels
    (i*stride + ph, j*stride + pw); tap (ky, kx) reaches it iff (ph + pad - ky*dil) % stride == 0 (same in x), reading
    dY at (i + (ph + pad - ky*dil)/stride, ...).  Returns [(ph, pw, [(dy, dx, kblk)])]; a phase without taps gets no
    gradient (the caller zero-fills)."""
    out = []
    for ph in range(stride):
        for pw in range(stride):
            taps = []
            for ky in range(kh):
                for kx in range(kw):
                    ny, nx = ph + pad - ky * dil, pw + pad - kx * dil
                    if ny % stride == 0 and nx % stride == 0:
                        taps.append((ny // stride, nx // stride, ky * kw + kx))
            out.append((ph, pw, taps))
    return out


class _ConvPacked(Function):
    """y = relu?(conv(x; wb) + bias (+ z)) with wb / wt the two bf16 GEMM packs of the SAME weight ([O, taps*I] and
    [I, taps*O]); the gradient w.r.t. the weight is returned for wb as fp32 [O, taps*I] (the layout the weight-gradient
    GEMM writes), wt gets none.

    Backward fusion flags (``opts``; all default off, every combination is exact):
      mask_input  x is the output of a ReLU whose backward masks by (x > 0) anyway: the input-gradient GEMM applies that
                  mask in its epilogue (the producer then only needs its bias column sums)
      premasked   the incoming gradient is already masked by (y > 0) (its producers all set mask_input): no mask pass
      x_sink      dict shared by the consumers of x that run on ONE stream: the first backward to run leaves its input
                  gradient there, the others add theirs into that tensor inside the GEMM epilogue and return none
      z_sink      the x_sink of the tensor passed as z: the identity-path gradient (returned to autograd as usual) is
                  registered there so that the other consumers of that tensor accumulate into it"""

    @staticmethod
    def forward(ctx, x, wb, wt, bias, z, kh, kw, stride, pad, dil, relu, holder, opts):
        ctx.holder = holder
        ctx.opts = dict(opts or {})
        if x.dtype != torch.bfloat16 or (z is not None and z.dtype != torch.bfloat16):
            # autograd would cast the returned gradient to the input's dtype (a copy): in-place sums would be lost
            ctx.opts.pop('x_sink', None)
            ctx.opts.pop('z_sink', None)
        x = G.as_nhwc(x, torch.bfloat16)
        B, C, H, W = x.shape
        O = wb.shape[0]
        assert C % 64 == 0 and O % 16 == 0 and wb.shape[1] == kh * kw * C
        Ho, Wo = conv_out_hw(H, W, kh, kw, stride, pad, dil)
        taps = [(ky * dil - pad, kx * dil - pad, ky * kw + kx) for ky in range(kh) for kx in range(kw)]
        if z is not None:
            z = G.as_nhwc(z.detach(), torch.bfloat16)
        y = conv2d_taps(x, wb.detach(), taps, Ho, Wo, stride, stride, bias=None if bias is None else bias.detach(),
                        resid=z, relu=relu)
        ctx.save_for_backward(x, wt, y if relu else None)
        ctx.cfg = (kh, kw, stride, pad, dil, relu, bias is not None, z is not None, O)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, wt, y = ctx.saved_tensors
        kh, kw, stride, pad, dil, relu, has_bias, has_z, O = ctx.cfg
        opts = ctx.opts
        B, C, H, W = x.shape
        Ho, Wo = gy.shape[2:]
        want_bias = has_bias and ctx.needs_input_grad[3]
        # premasked: the mask pass degenerates to the (read-only) column sums
        g, colsum = G.grad_prep(gy, y if (relu and not opts.get('premasked')) else None, want_bias)
        gx = gwb = gz = None
        if ctx.needs_input_grad[0]:
            sink = opts.get('x_sink')
            acc = sink.pop('dx', None) if sink is not None else None
            if acc is not None and not (acc.shape == (B, H, W, C) and acc.dtype == torch.bfloat16 and acc.is_contiguous()):
                sink['dx'] = acc
                acc = None
            phases = dgrad_phases(kh, kw, stride, pad, dil)
            full = all(len(t) for _, _, t in phases)
            buf = acc if acc is not None else (torch.empty if full else torch.zeros)((B, H, W, C), device=x.device,
                                                                                     dtype=torch.bfloat16)
            mask = x if opts.get('mask_input') else None
            for ph, pw, taps in phases:
                if not taps:
                    continue
                Hp, Wp = (H - ph + stride - 1) // stride, (W - pw + stride - 1) // stride
                if Hp <= 0 or Wp <= 0:
                    continue
                conv2d_taps(g, wt, taps, Hp, Wp, 1, 1, wt_taps=kh * kw, out=buf, out_map=(stride, stride, ph, pw, H, W),
                            resid=None if acc is None else acc.permute(0, 3, 1, 2), mask=mask)
            if sink is not None and acc is None:
                sink['dx'] = buf           # later consumers of x accumulate into this tensor
            gx = buf.permute(0, 3, 1, 2) if acc is None else None
        if ctx.needs_input_grad[1]:
            gwb = torch.zeros((O, kh * kw * C), device=x.device, dtype=torch.float32)
            _, _, _, _, ldy = G.nhwc_geom(g)
            L.call('lsnet_conv2d_wgrad_strided_nhwc_bf16', L.ptr(g), L.c_ll(ldy), L.ptr(x), L.c_ll(x.stride(3)),
                   L.c_int(B), L.c_int(H), L.c_int(W), L.c_int(C), L.c_int(Ho), L.c_int(Wo), L.c_int(O), L.c_int(kh),
                   L.c_int(kw), L.c_int(stride), L.c_int(stride), L.c_int(pad), L.c_int(pad), L.c_int(dil), L.c_int(dil),
                   L.ptr(gwb), L.stream())
        if gwb is not None and ctx.holder is not None:
            ctx.holder['gwb'] = gwb          # fp32, picked up by the fold's backward (see _BnFoldPacked)
            gwb = None
        if has_z and ctx.needs_input_grad[4]:
            gz = g
            zs = opts.get('z_sink')
            gb = g.permute(0, 2, 3, 1)
            if zs is not None and 'dx' not in zs and gb.is_contiguous() and g.dtype == torch.bfloat16:
                zs['dx'] = gb               # the identity gradient: the other consumers of z's tensor add theirs into it
        return gx, gwb, None, colsum, gz, None, None, None, None, None, None, None, None


def conv2d_packed(x, wb, wt, bias=None, z=None, kernel=(1, 1), stride=1, padding=0, dilation=1, relu=False,
                  wgrad_holder=None, opts=None):
    """``wgrad_holder``: dict that receives the fp32 weight gradient under 'gwb' instead of autograd (which would cast it
    to the bf16 pack's dtype); the producer of the packs reads it in its own backward.  ``opts``: see _ConvPacked."""
    return _ConvPacked.apply(x, wb, wt, bias, z, kernel[0], kernel[1], stride, padding, dilation, relu, wgrad_holder, opts)


class _ConvStrided(Function):
    """nn.Conv2d(stride = s) on the same kernels, straight from the fp32 OIHW parameter (FPN's stride-2 extra levels,
    mmdet/models/necks/fpn.py:203-211): packs are cached per parameter version."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, dil, relu):
        co, ci, kh, kw = weight.shape
        assert ci % 64 == 0 and co % 64 == 0, 'strided path: channel counts must be multiples of 64'
        x = G.as_nhwc(x, torch.bfloat16)
        B, _, H, W = x.shape
        wb = G.cached_pack(weight, 'sfwd', lambda t: t.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).to(torch.bfloat16).contiguous())
        b = None if bias is None else G.cached_pack(bias, 'bias%d' % co, lambda t: t.float().contiguous())
        Ho, Wo = conv_out_hw(H, W, kh, kw, stride, pad, dil)
        taps = [(ky * dil - pad, kx * dil - pad, ky * kw + kx) for ky in range(kh) for kx in range(kw)]
        y = conv2d_taps(x, wb, taps, Ho, Wo, stride, stride, bias=b, relu=relu)
        ctx.save_for_backward(x, weight, y if relu else None)
        ctx.cfg = (stride, pad, dil, relu, bias is not None)
        ctx.grad2d = G.direct_grad(weight)
        ctx.bias_param = bias
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, y = ctx.saved_tensors
        stride, pad, dil, relu, has_bias = ctx.cfg
        co, ci, kh, kw = weight.shape
        B, C, H, W = x.shape
        Ho, Wo = gy.shape[2:]
        g, colsum = G.grad_prep(gy, y if relu else None, has_bias and ctx.needs_input_grad[2],
                                colsum_into=G.direct_vec(ctx.bias_param))
        gx = gw = None
        if ctx.needs_input_grad[0]:
            wt = G.cached_pack(weight, 'sbwd', lambda t: t.permute(1, 2, 3, 0).reshape(ci, kh * kw * co).to(torch.bfloat16).contiguous())
            phases = dgrad_phases(kh, kw, stride, pad, dil)
            full = all(len(t) for _, _, t in phases)
            buf = (torch.empty if full else torch.zeros)((B, H, W, C), device=x.device, dtype=torch.bfloat16)
            for ph, pw, taps in phases:
                Hp, Wp = (H - ph + stride - 1) // stride, (W - pw + stride - 1) // stride
                if taps and Hp > 0 and Wp > 0:
                    conv2d_taps(g, wt, taps, Hp, Wp, 1, 1, wt_taps=kh * kw, out=buf, out_map=(stride, stride, ph, pw, H, W))
            gx = buf.permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            direct = ctx.grad2d
            if direct is not None and tuple(direct.shape) != (co, kh * kw * ci):
                direct = None
            dw = direct if direct is not None else torch.zeros((co, kh * kw * ci), device=x.device, dtype=torch.float32)
            L.call('lsnet_conv2d_wgrad_strided_nhwc_bf16', L.ptr(g), L.c_ll(G.nhwc_geom(g)[4]), L.ptr(x), L.c_ll(x.stride(3)),
                   L.c_int(B), L.c_int(H), L.c_int(W), L.c_int(C), L.c_int(Ho), L.c_int(Wo), L.c_int(co), L.c_int(kh),
                   L.c_int(kw), L.c_int(stride), L.c_int(stride), L.c_int(pad), L.c_int(pad), L.c_int(dil), L.c_int(dil),
                   L.ptr(dw), L.stream())
            if direct is None:
                gw = dw.view(co, kh, kw, ci).permute(0, 3, 1, 2).to(weight.dtype)
        return gx, gw, (colsum if has_bias and ctx.needs_input_grad[2] else None), None, None, None, None


def conv2d_strided(x, weight, bias=None, stride=2, padding=1, dilation=1, relu=False):
    return _ConvStrided.apply(x, weight, bias, stride, padding, dilation, relu)
