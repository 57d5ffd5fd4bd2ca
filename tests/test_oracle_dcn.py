"""The C restatement of the reference's DCN kernels (oracle/dcn_ref.c) against torchvision.ops.deform_conv2d (an
independent implementation of the same semantics) and fp64 gradcheck.  CPU only."""
import pytest
import torch
import torch.nn.functional as F
import torchvision

from oracle import dcn_ops as D


@pytest.mark.parametrize('groups,dg,stride,pad,dil', [(1, 1, 1, 1, 1), (2, 2, 2, 1, 1), (1, 1, 1, 2, 2), (4, 1, 2, 1, 1)])
def test_modulated_and_v1_vs_torchvision(groups, dg, stride, pad, dil):
    torch.manual_seed(0)
    dt = torch.float64
    B, C, H, W, Co = 2, 8, 7, 9, 8
    Ho = (H + 2 * pad - (dil * 2 + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * 2 + 1)) // stride + 1
    x = torch.randn(B, C, H, W, dtype=dt, requires_grad=True)
    off = (torch.randn(B, dg * 18, Ho, Wo, dtype=dt) * 2.5).requires_grad_()
    mask = torch.rand(B, dg * 9, Ho, Wo, dtype=dt, requires_grad=True)
    w = torch.randn(Co, C // groups, 3, 3, dtype=dt, requires_grad=True)
    b = torch.randn(Co, dtype=dt, requires_grad=True)
    y = D.modulated_deform_conv(x, off, mask, w, b, stride, pad, dil, groups, dg)
    y2 = torchvision.ops.deform_conv2d(x, off, w, b, stride=stride, padding=pad, dilation=dil, mask=mask)
    assert (y - y2).abs().max() < 1e-12
    g = torch.randn_like(y)
    for a, c in zip(torch.autograd.grad(y, [x, off, mask, w, b], g), torch.autograd.grad(y2, [x, off, mask, w, b], g)):
        assert (a - c).abs().max() < 1e-11
    y = D.deform_conv(x, off, w, stride, pad, dil, groups, dg)
    y2 = torchvision.ops.deform_conv2d(x, off, w, None, stride=stride, padding=pad, dilation=dil)
    assert (y - y2).abs().max() < 1e-12


def test_pyramid_gradcheck_and_grid_sample_formulation():
    torch.manual_seed(1)
    B, C, H, W, Co, Ho, Wo = 1, 4, 5, 6, 3, 3, 4
    x = torch.randn(B, C, H, W, dtype=torch.float64, requires_grad=True)
    off = (torch.randn(B, 18, Ho, Wo, dtype=torch.float64) * 1.3).requires_grad_()
    w = torch.randn(Co, C, 3, 3, dtype=torch.float64, requires_grad=True)
    sc = (H / Ho, W / Wo)
    assert torch.autograd.gradcheck(lambda a, b, c: D.pyramid_deform_conv(a, b, c, sc, 1, 1, 1, 1, 1), (x, off, w),
                                    eps=1e-6, atol=1e-5)
    y = D.pyramid_deform_conv(x, off, w, sc, 1, 1, 1)
    cols = []
    for k in range(9):
        i, j = k // 3, k % 3
        hs = (torch.arange(Ho, dtype=torch.float64).view(-1, 1) - 1 + i).float() * torch.tensor(sc[0]).float()
        ws = (torch.arange(Wo, dtype=torch.float64).view(1, -1) - 1 + j).float() * torch.tensor(sc[1]).float()
        hh, ww = hs.double() + off[:, 2 * k], ws.double() + off[:, 2 * k + 1]
        grid = torch.stack([2 * ww / (W - 1) - 1, 2 * hh / (H - 1) - 1], -1)
        cols.append(F.grid_sample(x, grid, mode='bilinear', padding_mode='zeros', align_corners=True))
    y2 = torch.einsum('ock,bckhw->bohw', w.view(Co, C, 9), torch.stack(cols, 2))
    assert (y - y2).abs().max() < 1e-12


def test_out_of_range_samples_are_zero_and_gradient_free():
    """P4: samples outside (-1, H) x (-1, W) contribute nothing, forward and backward."""
    x = torch.randn(1, 2, 4, 4, dtype=torch.float64, requires_grad=True)
    off = torch.full((1, 18, 4, 4), 50.0, dtype=torch.float64, requires_grad=True)
    w = torch.randn(3, 2, 3, 3, dtype=torch.float64)
    y = D.deform_conv(x, off, w, 1, 1, 1)
    assert float(y.abs().max()) == 0.0
    gx, go = torch.autograd.grad(y.sum(), [x, off])
    assert float(gx.abs().max()) == 0.0 and float(go.abs().max()) == 0.0
