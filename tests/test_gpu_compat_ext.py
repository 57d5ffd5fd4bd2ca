"""The operator-level drop-in executed: ``lsnet_b200.compat.deform_conv_ext`` and the reference's OWN
compiled extension (oracle/_ref, built from the unmodified sources) are called with IDENTICAL arguments through the
eight pybind entry points (mmdet/ops/dcn/src/deform_conv_ext.cpp:227-250), exactly as the reference's autograd Functions
call them (mmdet/ops/dcn/deform_conv.py:52-57, 88-103, 145-149, 163-170, 225-231, 252-277), and every output tensor is
compared.  Tolerance: the shim computes in bf16 with fp32 accumulation (inputs are rounded to bf16 first so that both
sides see the same operands): 1.5e-2 of each output's scale."""
import pytest
import torch

from oracle import build_ref

pytestmark = pytest.mark.gpu


def _ext():
    if build_ref.so_path() is None:
        pytest.skip('oracle/_ref not built (needs the reference tree at build time)')
    return build_ref.load_ext()


def _bf(t):
    return t.to(torch.bfloat16).float()


def _close(name, a, r, tol=1.5e-2):
    err = float((a - r).abs().max() / (r.abs().max() + 1e-30))
    assert err < tol, (name, err)


@pytest.mark.parametrize('B,C,Co,H,W', [(2, 64, 48, 13, 21), (2, 256, 256, 25, 42)])
def test_modulated_forward_backward_same_call(B, C, Co, H, W):
    from lsnet_b200.compat import deform_conv_ext as ours
    ref = _ext()
    g = torch.Generator().manual_seed(C + H)
    x = _bf(torch.randn(B, C, H, W, generator=g)).cuda()
    off = (torch.randn(B, 18, H, W, generator=g) * 1.5).cuda()
    mask = torch.rand(B, 9, H, W, generator=g).cuda()
    w = _bf(torch.randn(Co, C, 3, 3, generator=g) / (C * 9) ** 0.5).cuda()
    b = torch.randn(Co, generator=g).cuda()
    gy = _bf(torch.randn(B, Co, H, W, generator=g)).cuda()
    res = {}
    for name, ext in (('ref', ref), ('ours', ours)):
        out = x.new_empty(B, Co, H, W)
        e = x.new_empty(0)
        ext.modulated_deform_conv_forward(x, w, b, e, off, mask, out, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, True)
        gi, go, gm, gw, gb = (torch.zeros_like(t) for t in (x, off, mask, w, b))
        ext.modulated_deform_conv_backward(x, w, b, e, off, mask, e, gi, gw, gb, go, gm, gy, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, True)
        torch.cuda.synchronize()
        res[name] = (out, gi, go, gm, gw, gb)
    for n, a, r in zip(['out', 'grad_input', 'grad_offset', 'grad_mask', 'grad_weight', 'grad_bias'], res['ours'], res['ref']):
        _close(n, a, r)


def test_deform_v1_and_pyramid_same_call():
    from lsnet_b200.compat import deform_conv_ext as ours
    ref = _ext()
    g = torch.Generator().manual_seed(7)
    B, C, Co, H, W, Ho, Wo = 2, 64, 64, 25, 42, 13, 21
    x = _bf(torch.randn(B, C, H, W, generator=g)).cuda()
    w = _bf(torch.randn(Co, C, 3, 3, generator=g) / (C * 9) ** 0.5).cuda()
    # DCNv1 on the input grid
    off = (torch.randn(B, 18, H, W, generator=g) * 1.5).cuda()
    gy = _bf(torch.randn(B, Co, H, W, generator=g)).cuda()
    res = {}
    for name, ext in (('ref', ref), ('ours', ours)):
        e = x.new_empty(0)
        out = x.new_empty(B, Co, H, W)
        ext.deform_conv_forward(x, w, off, out, e, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 2)
        gi, go, gw = torch.zeros_like(x), torch.zeros_like(off), torch.zeros_like(w)
        ext.deform_conv_backward_input(x, off, gy, gi, go, w, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 2)
        ext.deform_conv_backward_parameters(x, off, gy, gw, e, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1.0, 2)
        torch.cuda.synchronize()
        res[name] = (out, gi, go, gw)
    for n, a, r in zip(['out', 'grad_input', 'grad_offset', 'grad_weight'], res['ours'], res['ref']):
        _close('v1 ' + n, a, r)
    # pyramid: sampling grid (Ho, Wo) on a finer map, non-dyadic scales
    offp = (torch.randn(B, 18, Ho, Wo, generator=g) * 1.5).cuda()
    gyp = _bf(torch.randn(B, Co, Ho, Wo, generator=g)).cuda()
    sh, sw = H / Ho, W / Wo
    for name, ext in (('ref', ref), ('ours', ours)):
        e = x.new_empty(0)
        out = x.new_empty(B, Co, Ho, Wo)
        ext.pyramid_deform_conv_forward(x, w, offp, out, e, e, 3, 3, 1, 1, 1, 1, 1, 1, sw, sh, 1, 1, 2)
        gi, go, gw = torch.zeros_like(x), torch.zeros_like(offp), torch.zeros_like(w)
        ext.pyramid_deform_conv_backward_input(x, offp, gyp, gi, go, w, e, 3, 3, 1, 1, 1, 1, 1, 1, sw, sh, 1, 1, 2)
        ext.pyramid_deform_conv_backward_parameters(x, offp, gyp, gw, e, e, 3, 3, 1, 1, 1, 1, 1, 1, sw, sh, 1, 1, 1.0, 2)
        torch.cuda.synchronize()
        res[name] = (out, gi, go, gw)
    for n, a, r in zip(['out', 'grad_input', 'grad_offset', 'grad_weight'], res['ours'], res['ref']):
        _close('pyramid ' + n, a, r)
