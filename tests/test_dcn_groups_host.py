"""Host logic of the grouped deformable convolution (no GPU): the block-diagonal expansion that lets the X-101 sites
(groups=64; mmdet/models/backbones/resnext.py:114-118) run on the dense kernels must be exact — a dense convolution with
the expanded weight equals the grouped convolution, and the block diagonal of the dense weight gradient (taken in the
kernels' [Cout, kh, kw, Cin] layout, as lsnet_b200/ops/dcn.py::_DCN.backward does) equals the grouped weight gradient."""
import pytest
import torch
import torch.nn.functional as F

from lsnet_b200.ops.dcn import _expand_groups


@pytest.mark.parametrize('groups,co,cig', [(4, 8, 3), (64, 128, 2), (1, 6, 5)])
def test_expand_groups_is_exact(groups, co, cig):
    torch.manual_seed(groups)
    ci = cig * groups
    w = torch.randn(co, cig, 3, 3, dtype=torch.float64)
    x = torch.randn(2, ci, 6, 7, dtype=torch.float64)
    wd = _expand_groups(w, groups)
    assert wd.shape == (co, ci, 3, 3)
    assert torch.allclose(F.conv2d(x, wd, padding=1), F.conv2d(x, w, padding=1, groups=groups), atol=1e-12)
    # zero outside the group's input block
    assert int((wd != 0).sum()) == w.numel()
    # tap-major (channels_last) parameters must expand the same way
    w_cl = w.contiguous(memory_format=torch.channels_last)
    assert torch.equal(_expand_groups(w_cl, groups), wd)
    # gradient: block diagonal of the dense dW in the kernels' [co, kh, kw, ci] layout
    wd_ = wd.clone().requires_grad_(True)
    wg_ = w.clone().requires_grad_(True)
    F.conv2d(x, wd_, padding=1).square().sum().backward()
    F.conv2d(x, wg_, padding=1, groups=groups).square().sum().backward()
    d = wd_.grad.permute(0, 2, 3, 1).reshape(co, 3, 3, ci)
    if groups > 1:
        idx = torch.arange(groups)
        d = d.view(groups, co // groups, 3, 3, groups, cig)[idx, :, :, :, idx].reshape(co, 3, 3, cig)
    assert torch.allclose(d.permute(0, 3, 1, 2), wg_.grad, atol=1e-10)
