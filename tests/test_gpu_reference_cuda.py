"""Pins the oracle's DCN restatement (oracle/dcn_ref.c, CPU) to the REFERENCE'S OWN CUDA kernels: the unmodified
extension compiled into oracle/_ref (oracle/build_ref.py) is driven through its 8 pybind entry points
(mmdet/ops/dcn/src/deform_conv_ext.cpp:227-250) exactly as the reference's autograd Functions call them
(mmdet/ops/dcn/deform_conv.py:52-57,88-103,145-149,163-170,225-231,252-277), in fp32, and compared with the oracle on
the same inputs.  (fp64 is not usable: the reference launches 1024-thread blocks unconditionally, the fp64 col2im
instantiation needs more registers than that allows on sm_100, and the reference only printf's launch errors
(deform_conv_cuda_kernel.cu:326-330), leaving grad_input silently zero.)  Tolerance 2e-4 of the output scale: fp32
atomics order + the FMA contraction nvcc applies to the sampling position.  Also times the reference kernel against ours on a head-sized shape ("kernel to beat")."""
import pytest
import torch

from oracle import build_ref
from oracle import dcn_ops as OD

pytestmark = pytest.mark.gpu


def _ext():
    if build_ref.so_path() is None:
        pytest.skip('oracle/_ref not built (needs the reference tree at build time)')
    return build_ref.load_ext()


def _ref_modulated(ext, x, off, mask, w, b, gy):
    out = x.new_empty(gy.shape)
    e = x.new_empty(0)
    ext.modulated_deform_conv_forward(x, w, b, e, off, mask, out, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, True)
    gi, go, gm, gw, gb = (torch.zeros_like(t) for t in (x, off, mask, w, b))
    ext.modulated_deform_conv_backward(x, w, b, e, off, mask, e, gi, gw, gb, go, gm, gy.contiguous(), 3, 3, 1, 1, 1, 1, 1, 1,
                                       1, 1, True)
    return out, gi, go, gm, gw, gb


def test_oracle_dcnv2_equals_reference_cuda():
    ext = _ext()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 9, 11, generator=g, dtype=torch.float32)
    off = torch.randn(2, 18, 9, 11, generator=g, dtype=torch.float32) * 2.5
    mask = torch.rand(2, 9, 9, 11, generator=g, dtype=torch.float32)
    w = torch.randn(6, 8, 3, 3, generator=g, dtype=torch.float32)
    b = torch.randn(6, generator=g, dtype=torch.float32)
    gy = torch.randn(2, 6, 9, 11, generator=g, dtype=torch.float32)
    ref = _ref_modulated(ext, *(t.cuda() for t in (x, off, mask, w, b, gy)))
    xr, offr, mr, wr, br = (t.clone().requires_grad_(True) for t in (x, off, mask, w, b))
    y = OD.modulated_deform_conv(xr, offr, mr, wr, br, 1, 1, 1)
    grads = torch.autograd.grad(y, [xr, offr, mr, wr, br], gy)
    for name, a, r in zip(['out', 'dx', 'doff', 'dmask', 'dw', 'db'], (y,) + grads, ref):
        assert (a.detach() - r.cpu()).abs().max() < 2e-4 * max(1.0, float(r.abs().max())), name


def test_oracle_pyramid_equals_reference_cuda():
    ext = _ext()
    g = torch.Generator().manual_seed(1)
    H, W, Ho, Wo = 13, 21, 7, 11
    x = torch.randn(2, 8, H, W, generator=g, dtype=torch.float32)
    off = torch.randn(2, 18, Ho, Wo, generator=g, dtype=torch.float32) * 1.5
    w = torch.randn(6, 8, 3, 3, generator=g, dtype=torch.float32)
    gy = torch.randn(2, 6, Ho, Wo, generator=g, dtype=torch.float32)
    sh, sw = H / Ho, W / Wo
    xc, oc, wc, gc = (t.cuda() for t in (x, off, w, gy))
    e = xc.new_empty(0)
    out = xc.new_empty(gy.shape)
    ext.pyramid_deform_conv_forward(xc, wc, oc, out, e, e, 3, 3, 1, 1, 1, 1, 1, 1, sw, sh, 1, 1, 2)
    gi, go, gw = torch.zeros_like(xc), torch.zeros_like(oc), torch.zeros_like(wc)
    ext.pyramid_deform_conv_backward_input(xc, oc, gc, gi, go, wc, e, 3, 3, 1, 1, 1, 1, 1, 1, sw, sh, 1, 1, 2)
    ext.pyramid_deform_conv_backward_parameters(xc, oc, gc, gw, e, e, 3, 3, 1, 1, 1, 1, 1, 1, sw, sh, 1, 1, 1, 2)
    xr, offr, wr = (t.clone().requires_grad_(True) for t in (x, off, w))
    y = OD.pyramid_deform_conv(xr, offr, wr, (sh, sw), 1, 1, 1)
    grads = torch.autograd.grad(y, [xr, offr, wr], gy)
    # the scale is a C float in the reference; the oracle applies the same float product (dcn_ref.c sample_pos)
    for name, a, r in zip(['out', 'dx', 'doff', 'dw'], (y,) + grads, (out, gi, go, gw)):
        assert (a.detach() - r.cpu()).abs().max() < 2e-4 * max(1.0, float(r.abs().max())), name


def test_reference_kernel_vs_ours_timing():
    """Head level-1 shape (B=4, 256 ch, 50x84): reference fp32 kernel vs the B200 path, forward only; informational
    (printed with -s / -rA), asserts only that ours is not slower."""
    import lsnet_b200.ops as ops
    ext = _ext()
    g = torch.Generator().manual_seed(2)
    B, C, H, W = 4, 256, 50, 84
    x = torch.randn(B, C, H, W, generator=g).cuda()
    off = (torch.randn(B, 18, H, W, generator=g) * 1.5).cuda()
    mask = torch.rand(B, 9, H, W, generator=g).cuda()
    w = (torch.randn(256, 256, 3, 3, generator=g) / 48).cuda()
    b = torch.zeros(256).cuda()
    out = torch.empty(B, 256, H, W, device='cuda')
    e = x.new_empty(0)

    def t(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t_ref = t(lambda: ext.modulated_deform_conv_forward(x, w, b, e, off, mask, out, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, True))
    xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    t_our = t(lambda: ops.modulated_deform_conv(xb, off, mask, w, b, 1, 1, 1))
    print(f'DCNv2 fwd B4 C256 50x84: reference kernel {t_ref:.3f} ms, lsnet_b200 {t_our:.3f} ms, x{t_ref / t_our:.1f}')
    assert t_our < t_ref
