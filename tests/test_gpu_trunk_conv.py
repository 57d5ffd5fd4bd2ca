"""GPU parity tests of the trunk convolution path on the library's own tcgen05 kernels (SURVEY §8 rows a11 / f4 / a4's
stride-2 levels): the general strided implicit-GEMM convolution (forward with bias + identity add + ReLU epilogue, input
gradient by output phases, strided weight gradient) and the BN fold that produces its operand packs, against plain fp32
PyTorch convolutions on the same bf16-rounded operands (the reference computes them with nn.Conv2d / nn.BatchNorm2d,
mmdet/models/backbones/resnet.py:261-301, 636-646).

Tolerances: bf16 operands, fp32 accumulation on both sides; outputs and input gradients are stored in bf16 (one 2^-9
rounding: bound 1.5e-2 of the output scale), weight gradients in fp32 (summation order only: 4e-3)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _bf(t):
    return t.to(torch.bfloat16).float()


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _packs(w):
    """(O, I, kh, kw) fp32 -> wb bf16 [O, taps*I], wt bf16 [I, taps*O]."""
    O, I, kh, kw = w.shape
    wb = w.permute(0, 2, 3, 1).reshape(O, kh * kw * I).to(torch.bfloat16).contiguous()
    wt = w.permute(1, 2, 3, 0).reshape(I, kh * kw * O).to(torch.bfloat16).contiguous()
    return wb, wt


CASES = [  # B, I, O, H, W, k, stride, pad, dil, with_z, relu
    (2, 64, 64, 40, 56, 1, 1, 0, 1, False, True),
    (2, 64, 256, 25, 42, 1, 1, 0, 1, True, True),
    (2, 128, 128, 25, 42, 3, 1, 1, 1, False, True),
    (2, 128, 128, 26, 43, 3, 2, 1, 1, False, True),       # odd extents: ragged phases
    (2, 256, 512, 25, 42, 1, 2, 0, 1, False, False),      # downsample branch: no bias ReLU
    (1, 256, 256, 13, 21, 3, 2, 1, 1, False, False),      # FPN extra level 13x21 -> 7x11
    (4, 512, 512, 50, 84, 3, 2, 1, 1, False, True),
    (2, 64, 64, 33, 47, 3, 1, 2, 2, False, True),         # dilation 2
]


@pytest.mark.parametrize('B,I,O,H,W,k,stride,pad,dil,with_z,relu', CASES)
def test_conv2d_packed_matches_torch(B, I, O, H, W, k, stride, pad, dil, with_z, relu):
    from lsnet_b200.ops.conv import conv2d_packed, conv_out_hw
    g = torch.Generator().manual_seed(B * 1000 + I + O + H + W + k + stride)
    x = _bf(torch.randn(B, I, H, W, generator=g)).to(DEV)
    w = _bf(torch.randn(O, I, k, k, generator=g) / (I * k * k) ** 0.5).to(DEV)
    bias = torch.randn(O, generator=g).to(DEV)
    Ho, Wo = conv_out_hw(H, W, k, k, stride, pad, dil)
    z = _bf(torch.randn(B, O, Ho, Wo, generator=g)).to(DEV) if with_z else None
    gy = _bf(torch.randn(B, O, Ho, Wo, generator=g)).to(DEV)

    # reference: fp32 conv on the same bf16-rounded operands
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    zr = z.clone().requires_grad_(True) if with_z else None
    yr = F.conv2d(xr, wr, br, stride, pad, dil)
    if with_z:
        yr = yr + zr
    if relu:
        yr = F.relu(yr)
    yr.backward(gy)

    xo = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wb, wt = _packs(w)
    holder = {}
    wb.requires_grad_(True)
    bo = bias.clone().requires_grad_(True)
    zo = z.to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True) if with_z else None
    yo = conv2d_packed(xo, wb, wt, bo, zo, (k, k), stride, pad, dil, relu, wgrad_holder=holder)
    assert yo.shape == yr.shape
    assert _rel(yo.float(), yr.detach()) < 1.5e-2
    # the ReLU mask must come from the SAME forward values: compare gradients where both sides agree on the mask
    yo.backward(gy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
    torch.cuda.synchronize()
    same = ((yo.float() > 0) == (yr > 0)) if relu else torch.ones_like(yr, dtype=torch.bool)
    assert float(same.float().mean()) > 0.995
    assert _rel(xo.grad.float(), xr.grad) < 2.5e-2
    gwb = holder['gwb'].view(O, k, k, I).permute(0, 3, 1, 2)
    assert _rel(gwb, wr.grad) < 2e-2          # includes the <= 0.5 % of positions whose ReLU mask differs (bf16 forward)
    assert _rel(bo.grad, br.grad) < 1e-2
    if with_z:
        assert _rel(zo.grad.float(), zr.grad) < 2.5e-2


@pytest.mark.parametrize('tap_major', [False, True])
@pytest.mark.parametrize('O,I,k', [(64, 64, 1), (128, 64, 3), (96, 192, 3)])
def test_bn_fold_packed(O, I, k, tap_major):
    from lsnet_b200.modules.backbone import _BnFoldPacked
    g = torch.Generator().manual_seed(O + I + k)
    W = torch.randn(O, I, k, k, generator=g).to(DEV)
    if tap_major:
        W = W.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    gamma = (torch.rand(O, generator=g) + 0.5).to(DEV)
    beta = torch.randn(O, generator=g).to(DEV)
    mean = torch.randn(O, generator=g).to(DEV) * 0.1
    var = (torch.rand(O, generator=g) + 0.5).to(DEV)
    eps = 1e-5
    gwb = torch.randn(O, k * k * I, generator=g).to(DEV)
    gbias = torch.randn(O, generator=g).to(DEV)

    Wr, gr, br = W.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    s = gr / torch.sqrt(var + eps)
    wbr = (Wr * s.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).reshape(O, -1)
    biasr = br - mean * s
    ((wbr * gwb).sum() + (biasr * gbias).sum()).backward()

    Wo, go, bo = W.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    holder = {}
    wb, wt, bias = _BnFoldPacked.apply(Wo, go, bo, mean, var, eps, holder)
    assert _rel(wb.float(), wbr.detach()) < 8e-3 and _rel(bias, biasr.detach()) < 1e-6
    assert torch.equal(wt.view(I, k * k, O).permute(2, 1, 0).reshape(O, -1), wb)
    holder['gwb'] = gwb
    (bias * gbias).sum().backward()        # the weight gradient arrives through the holder
    torch.cuda.synchronize()
    assert _rel(Wo.grad, Wr.grad) < 1e-5 and _rel(go.grad, gr.grad) < 1e-4 and _rel(bo.grad, br.grad) < 1e-6


def test_backbone_own_convs_match_cudnn_path():
    """ResNet-50 trunk: the library's own convolutions vs the cuDNN fused path on identical parameters -- outputs of the
    four stages and every parameter gradient (both paths compute in bf16 with fp32 accumulation)."""
    import lsnet_b200 as L
    from lsnet_b200.modules import backbone as bb
    torch.manual_seed(0)
    net = L.build_backbone(dict(type='ResNet', depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1,
                                norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, style='pytorch'))
    net.init_weights(None)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
    net.cuda().train()
    x = torch.randn(2, 3, 256, 320, device='cuda').contiguous(memory_format=torch.channels_last)

    def run(mode):
        old = bb.TRUNK
        bb.TRUNK = mode
        try:
            net.zero_grad()
            outs = net(x)
            loss = sum((o.float() ** 2).mean() for o in outs)
            loss.backward()
            torch.cuda.synchronize()
        finally:
            bb.TRUNK = old
        return [o.float().detach() for o in outs], {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    o1, g1 = run('own')
    o0, g0 = run('cudnn')
    for a, b in zip(o1, o0):
        assert _rel(a, b) < 4e-2
    assert set(g1) == set(g0)
    for k in g0:
        if g0[k].norm() > 0:
            cos = float(torch.dot(g1[k].flatten().float(), g0[k].flatten().float()) / (g1[k].norm() * g0[k].norm()))
            assert cos > 0.99, (k, cos)


@pytest.mark.parametrize('B,I,O,H,W', [(2, 256, 256, 13, 21), (2, 2048, 256, 25, 42), (1, 64, 64, 7, 11)])
def test_conv2d_strided_from_parameter(B, I, O, H, W):
    """FPN's stride-2 3x3 extra-level conv straight from the fp32 OIHW parameter (necks/fpn.py:203-211)."""
    from lsnet_b200.ops import conv2d_strided
    g = torch.Generator().manual_seed(I + O + H)
    x = _bf(torch.randn(B, I, H, W, generator=g)).to(DEV)
    w = _bf(torch.randn(O, I, 3, 3, generator=g) / (I * 9) ** 0.5).to(DEV)
    bias = torch.randn(O, generator=g).to(DEV)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, 2, 1)
    gy = _bf(torch.randn(yr.shape, generator=g)).to(DEV)
    yr.backward(gy)
    xo = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wo, bo = w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    yo = conv2d_strided(xo, wo, bo, 2, 1)
    assert _rel(yo.float(), yr.detach()) < 1.5e-2
    yo.backward(gy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
    torch.cuda.synchronize()
    assert _rel(xo.grad.float(), xr.grad) < 2.5e-2
    assert _rel(wo.grad, wr.grad) < 1e-2 and _rel(bo.grad, br.grad) < 1e-2


@pytest.mark.parametrize('B,H,W,nhwc,dtype', [(2, 64, 96, False, torch.float32), (1, 75, 133, False, torch.float32),
                                              (2, 50, 70, True, torch.float32), (1, 64, 64, False, torch.bfloat16)])
def test_stem_conv_and_maxpool(B, H, W, nhwc, dtype):
    """conv1 7x7/2 + frozen BN + ReLU (tcgen05 implicit GEMM with an in-kernel A tile) and the 3x3/2 max-pool vs PyTorch
    (resnet.py:619-623); ragged sizes exercise the zero padding and partial tiles."""
    from lsnet_b200 import ops
    g = torch.Generator().manual_seed(H * W)
    x = _bf(torch.randn(B, 3, H, W, generator=g)).to(DEV)
    w = torch.randn(64, 3, 7, 7, generator=g).to(DEV) / 12
    gamma, beta = (torch.rand(64, generator=g) + 0.5).to(DEV), torch.randn(64, generator=g).to(DEV) * 0.1
    mean, var = torch.randn(64, generator=g).to(DEV) * 0.1, (torch.rand(64, generator=g) + 0.5).to(DEV)
    wp, shift = ops.pack_stem_weight(w, gamma, beta, mean, var, 1e-5)
    s = gamma / torch.sqrt(var + 1e-5)
    wf = _bf(w * s.view(-1, 1, 1, 1))
    ref = F.relu(F.conv2d(x, wf, shift, 2, 3))
    xi = x.to(dtype)
    if nhwc:
        xi = xi.contiguous(memory_format=torch.channels_last)
    y = ops.stem_conv(xi, wp, shift)
    torch.cuda.synchronize()
    assert y.shape == ref.shape
    assert _rel(y.float(), ref) < 1.5e-2
    pooled = ops.maxpool3x3s2(y)
    refp = F.max_pool2d(y.float(), 3, 2, 1)
    assert pooled.shape == refp.shape and torch.equal(pooled.float(), refp)


def test_backbone_backward_fusion_is_exact():
    """The trunk's fused backward (ReLU masks and identity-path sums inside the input-gradient GEMM epilogues, shared
    gradient sinks) against the plain per-op backward of the SAME kernels: identical math, so outputs are bit-equal and
    gradients agree to bf16 rounding of the partial sums."""
    import lsnet_b200 as L
    from lsnet_b200.modules import backbone as bb
    torch.manual_seed(1)
    net = L.build_backbone(dict(type='ResNet', depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1,
                                norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, style='pytorch'))
    net.init_weights(None)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.1)
    net.cuda().train()
    x = torch.randn(2, 3, 200, 264, device='cuda')

    def run(fuse):
        old = bb.BWD_FUSE
        bb.BWD_FUSE = fuse
        try:
            net.zero_grad()
            outs = net(x)
            loss = sum((o.float() ** 2).mean() for o in outs)
            loss.backward()
            torch.cuda.synchronize()
        finally:
            bb.BWD_FUSE = old
        return [o.float().detach() for o in outs], {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    o1, g1 = run(True)
    o0, g0 = run(False)
    for a, b in zip(o1, o0):
        assert torch.equal(a, b)
    assert set(g1) == set(g0)
    worst = 0.0
    for k in g0:
        n0 = float(g0[k].norm())
        if n0 > 0:
            worst = max(worst, float((g1[k] - g0[k]).norm()) / n0)
    assert worst < 2e-2, worst
