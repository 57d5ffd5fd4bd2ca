"""Fused DCN kernels (lsnet_b200/csrc/dcn_fused.cu) through the whole-operator C ABI (lsnet_dcn_forward /
lsnet_dcn_backward_data / lsnet_dcn_backward_weight):

* against the column-matrix path of the same library on identical inputs (only the fp32 accumulation order inside the
  tensor core differs: <= 1e-3 of the output scale; the column side output is BIT-exact with lsnet_dcn_im2col_bf16),
* against the oracle (oracle/dcn_ref.c, the CPU restatement of the reference kernels) for DCNv2 / DCNv1 / pyramid DCN
  at shapes that exercise every tile edge: ragged patches, several channel blocks, N below / at the 256-wide tile,
  strides, large offsets (samples outside the map), non-dyadic pyramid scales,
* at the BASELINE level shapes (B=4, C=256, 100x168 ... 7x11) against the REFERENCE'S OWN CUDA kernels (oracle/_ref,
  fp32) with an explicit bf16 error budget per output.
"""
import pytest
import torch

from oracle import dcn_ops as OD

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _ops():
    import lsnet_b200.ops as ops
    return ops


def _lib():
    from lsnet_b200 import lib
    return lib.load()


def _bf(t):
    return t.to(torch.bfloat16).float()


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _nhwc(t, dtype=None):
    t = t.to(DEV)
    if dtype is not None:
        t = t.to(dtype)
    return t.contiguous(memory_format=torch.channels_last)


def _pack(w):
    co, ci, kh, kw = w.shape
    npad = (co + 15) // 16 * 16
    p = torch.zeros(npad, kh * kw * ci, dtype=torch.bfloat16, device=DEV)
    p[:co] = w.permute(0, 2, 3, 1).reshape(co, -1).to(DEV, torch.bfloat16)
    return p


CASES = [
    # name,        B, C,   H,  W,  Ho, Wo, Co,  stride, mag, mask
    ('v2_small', 2, 64, 13, 21, 13, 21, 48, 1, 2.5, True),
    ('v2_c256_n256', 1, 256, 25, 42, 25, 42, 256, 1, 1.5, True),
    ('v2_far', 2, 128, 9, 10, 9, 10, 256, 1, 12.0, True),       # most samples fall outside the map
    ('v1_n32', 2, 64, 17, 19, 17, 19, 32, 1, 2.0, False),
    ('v2_stride2', 2, 64, 24, 30, 12, 15, 64, 2, 1.5, True),
    ('pyr_up', 2, 64, 13, 21, 25, 42, 64, 1, 2.5, False),        # output grid finer than the sampled map (non-dyadic)
    ('pyr_down', 2, 128, 25, 42, 13, 21, 128, 1, 2.5, False),
    ('pyr_tiny', 4, 256, 13, 21, 7, 11, 256, 1, 1.0, False),
]


def _make(case):
    name, B, C, H, W, Ho, Wo, Co, stride, mag, has_mask = case
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    x = _bf(torch.randn(B, C, H, W, generator=g))
    off = torch.randn(B, 18, Ho, Wo, generator=g) * mag
    mask = torch.rand(B, 9, Ho, Wo, generator=g) if has_mask else None
    w = _bf(torch.randn(Co, C, 3, 3, generator=g) / (C * 9) ** 0.5)
    bias = torch.randn(Co, generator=g) if has_mask else None
    return x, off, mask, w, bias


def _geom(case):
    name, B, C, H, W, Ho, Wo, Co, stride, mag, has_mask = case
    pyr = name.startswith('pyr')
    scales = (H / Ho, W / Wo) if pyr else (1.0, 1.0)
    return (Ho, Wo, 3, 3, (stride, stride), (1, 1), (1, 1), scales, 1)


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_fused_forward_vs_unfused_and_oracle(case):
    ops, lib = _ops(), _lib()
    name, B, C, H, W, Ho, Wo, Co, stride, mag, has_mask = case
    x, off, mask, w, bias = _make(case)
    cfg = _geom(case)
    xd, offd = _nhwc(x, torch.bfloat16), _nhwc(off)
    maskd = _nhwc(mask) if mask is not None else None
    wp = _pack(w)
    bd = None
    if bias is not None:
        bd = torch.zeros(wp.shape[0], device=DEV)
        bd[:Co] = bias.to(DEV)
    try:
        lib.lsnet_dcn_fused_enable(1)
        out_f, col_f = ops.dcn_forward(xd, offd, maskd, wp, bd, *cfg, out_dtype=torch.float32, save_col=True)
        out_f2, _ = ops.dcn_forward(xd, offd, maskd, wp, bd, *cfg, out_dtype=torch.float32, save_col=False)
        out_b, _ = ops.dcn_forward(xd, offd, maskd, wp, bd, *cfg, out_dtype=torch.bfloat16, relu=True)
        lib.lsnet_dcn_fused_enable(0)
        out_u, col_u = ops.dcn_forward(xd, offd, maskd, wp, bd, *cfg, out_dtype=torch.float32, save_col=True)
    finally:
        lib.lsnet_dcn_fused_enable(2)
    torch.cuda.synchronize()
    # the side output is the same arithmetic as the standalone gather: bit-exact
    assert torch.equal(col_f, col_u), name
    assert torch.equal(out_f, out_f2), name
    assert _rel(out_f, out_u) < 1e-3, (name, _rel(out_f, out_u))
    assert _rel(out_b.float(), out_u.clamp(min=0)) < 1.5e-2, name
    # oracle
    sc = cfg[7]
    if name.startswith('pyr'):
        ref = OD.pyramid_deform_conv(x, off, w, sc, stride, 1, 1)
    elif mask is None:
        ref = OD.deform_conv(x, off, w, stride, 1, 1)
    else:
        ref = OD.modulated_deform_conv(x, off, mask, w, bias, stride, 1, 1)
    got = out_f.view(B, Ho, Wo, -1)[..., :Co].permute(0, 3, 1, 2)
    assert got.shape == ref.shape
    assert _rel(got, ref) < 4e-3, (name, _rel(got, ref))


def test_fused_forward_into_channel_slice_and_logits():
    """The three pyramid DCNs of a level write into channel slices of ONE [B,H,W,768] buffer (LSHead); DCNv2 reads the
    mask LOGITS from the conv_offset output."""
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    B, C, H, W = 2, 64, 11, 13
    x = _nhwc(_bf(torch.randn(B, C, H, W, generator=g)), torch.bfloat16)
    om = _nhwc(torch.randn(B, 32, H, W, generator=g))          # conv_offset output padded to 32 channels
    w = _bf(torch.randn(64, C, 3, 3, generator=g) / 24)
    wp = _pack(w)
    cfg = (H, W, 3, 3, (1, 1), (1, 1), (1, 1), (1.0, 1.0), 1)
    buf = torch.full((B, H, W, 192), 7.0, device=DEV, dtype=torch.bfloat16)
    out2d = torch.as_strided(buf, (B * H * W, 64), (192, 1), 64)
    ops.dcn_forward(x, om[:, :18], om[:, 18:27], wp, None, *cfg, mask_logits=True, out=out2d)
    ref, _ = ops.dcn_forward(x, om[:, :18], torch.sigmoid(om[:, 18:27]), wp, None, *cfg, mask_logits=False)
    torch.cuda.synchronize()
    assert _rel(buf[..., 64:128].reshape(-1, 64).float(), ref.float()) < 1.5e-2
    assert float((buf[..., :64] - 7).abs().max()) == 0 and float((buf[..., 128:] - 7).abs().max()) == 0


LEVELS = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]


@pytest.fixture(params=['fused', 'columns'])
def dcn_path(request):
    """Run the test once with the fused kernels forced on every supported shape and once on the column-matrix path
    (the library's default picks per shape)."""
    lib = _lib()
    lib.lsnet_dcn_fused_enable(1 if request.param == 'fused' else 0)
    yield request.param
    lib.lsnet_dcn_fused_enable(2)


def _ref_ext():
    from oracle import build_ref
    if build_ref.so_path() is None:
        pytest.skip('oracle/_ref not built (needs the reference tree at build time)')
    return build_ref.load_ext()


@pytest.mark.parametrize('lvl', range(5))
def test_dcnv2_baseline_shape_vs_reference_cuda(lvl, dcn_path):
    """B=4, C=256 -> 256 on the five level grids of the 800x1344 BASELINE config: our bf16 operator (forward + all
    gradients, through ops.modulated_deform_conv -> the whole-operator C ABI) against the reference's own fp32 CUDA
    kernels on bf16-representable inputs.  Error budget (relative to the max |reference| of each output):
      out     4e-3  columns rounded to bf16 (2^-9 per element, averaged over K = 2304 products)
      dW      4e-3  bf16 columns and dY, fp32 accumulation over up to 67 200 pixels
      dOffset 2e-2  dCol rounded to bf16 before the <dCol, x> reductions over 256 channels
      dMask   2e-2  same
      dX      4e-2  bf16 dCol + packed-bf16 atomic accumulation (default) -- 2e-2 with fp32 accumulation
      db      1e-3  fp32 column sums of bf16 dY"""
    ops = _ops()
    ext = _ref_ext()
    H, W = LEVELS[lvl]
    B, C, Co = 4, 256, 256
    g = torch.Generator().manual_seed(100 + lvl)
    x = _bf(torch.randn(B, C, H, W, generator=g)).to(DEV)
    off = (torch.randn(B, 18, H, W, generator=g) * 1.5).to(DEV)
    mask = torch.rand(B, 9, H, W, generator=g).to(DEV)
    w = _bf(torch.randn(Co, C, 3, 3, generator=g) / (C * 9) ** 0.5).to(DEV)
    b = torch.randn(Co, generator=g).to(DEV)
    gy = _bf(torch.randn(B, Co, H, W, generator=g)).to(DEV)
    e = x.new_empty(0)
    ref_out = torch.empty(B, Co, H, W, device=DEV)
    ext.modulated_deform_conv_forward(x, w, b, e, off, mask, ref_out, e, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, True)
    gi, go, gm, gw, gb = (torch.zeros_like(t) for t in (x, off, mask, w, b))
    ext.modulated_deform_conv_backward(x, w, b, e, off, mask, e, gi, gw, gb, go, gm, gy.contiguous(), 3, 3, 1, 1, 1, 1, 1,
                                       1, 1, 1, True)
    ins = [t.clone().requires_grad_(True) for t in (x, off, mask, w, b)]
    out = ops.modulated_deform_conv(*ins, 1, 1, 1, out_fp32=True)
    grads = torch.autograd.grad(out, ins, gy)
    torch.cuda.synchronize()
    budget = dict(out=4e-3, dx=4e-2, doff=2e-2, dmask=2e-2, dw=4e-3, db=1e-3)
    got = dict(out=out, dx=grads[0], doff=grads[1], dmask=grads[2], dw=grads[3], db=grads[4])
    ref = dict(out=ref_out, dx=gi, doff=go, dmask=gm, dw=gw, db=gb)
    errs = {k: _rel(got[k].float(), ref[k]) for k in budget}
    print(f'level {lvl} ({H}x{W}) rel err vs reference CUDA:', {k: f'{v:.2e}' for k, v in errs.items()})
    for k, tol in budget.items():
        assert errs[k] < tol, (lvl, k, errs[k])


@pytest.mark.parametrize('lvl,src', [(0, 1), (1, 0), (2, 3), (4, 2)])
def test_pyramid_baseline_shape_vs_reference_cuda(lvl, src, dcn_path):
    """Pyramid DCN on the BASELINE level grids: output/offset grid = level `lvl`, sampled map = level `src` (scale_h =
    H_src / H_lvl, non-dyadic for 25->13, 13->7), against the reference's pyramid kernels (fp32)."""
    ops = _ops()
    ext = _ref_ext()
    (Ho, Wo), (H, W) = LEVELS[lvl], LEVELS[src]
    B, C, Co = 4, 256, 256
    sh, sw = H / Ho, W / Wo
    g = torch.Generator().manual_seed(200 + 10 * lvl + src)
    x = _bf(torch.randn(B, C, H, W, generator=g)).to(DEV)
    off = (torch.randn(B, 18, Ho, Wo, generator=g) * 2.0).to(DEV)
    w = _bf(torch.randn(Co, C, 3, 3, generator=g) / (C * 9) ** 0.5).to(DEV)
    gy = _bf(torch.randn(B, Co, Ho, Wo, generator=g)).to(DEV)
    e = x.new_empty(0)
    step = min(64, B)
    ref_out = torch.empty(B, Co, Ho, Wo, device=DEV)
    ext.pyramid_deform_conv_forward(x, w, off, ref_out, e, e, 3, 3, 1, 1, 1, 1, 1, 1, sw, sh, 1, 1, step)
    gi, go, gw = torch.zeros_like(x), torch.zeros_like(off), torch.zeros_like(w)
    ext.pyramid_deform_conv_backward_input(x, off, gy.contiguous(), gi, go, w, e, 3, 3, 1, 1, 1, 1, 1, 1, sw, sh, 1, 1, step)
    ext.pyramid_deform_conv_backward_parameters(x, off, gy.contiguous(), gw, e, e, 3, 3, 1, 1, 1, 1, 1, 1, sw, sh, 1, 1, 1,
                                                step)
    ins = [t.clone().requires_grad_(True) for t in (x, off, w)]
    out = ops.pyramid_deform_conv(ins[0], ins[1], ins[2], (sh, sw), 1, 1, 1, out_fp32=True)
    grads = torch.autograd.grad(out, ins, gy)
    torch.cuda.synchronize()
    # dX: up to (sh*sw)^-1 * 36 contributions per input pixel accumulate in packed bf16 (see test_gpu_kernels.py)
    budget = dict(out=4e-3, dx=6e-2 if sh < 1 else 4e-2, doff=2e-2, dw=4e-3)
    got = dict(out=out, dx=grads[0], doff=grads[1], dw=grads[2])
    ref = dict(out=ref_out, dx=gi, doff=go, dw=gw)
    errs = {k: _rel(got[k].float(), ref[k]) for k in budget}
    print(f'pyramid {Ho}x{Wo} <- {H}x{W} rel err vs reference CUDA:', {k: f'{v:.2e}' for k, v in errs.items()})
    for k, tol in budget.items():
        assert errs[k] < tol, (lvl, src, k, errs[k])


WG_CASES = [
    # name,      B, C,   H,  W,  Ho, Wo, M,  mag, mask
    ('wg_v2', 2, 256, 13, 21, 13, 21, 256, 2.0, True),
    ('wg_v2_ragged_m48', 3, 256, 9, 19, 9, 19, 48, 2.0, True),
    ('wg_c512_m320', 1, 512, 10, 12, 10, 12, 320, 1.5, True),      # 2 channel tiles, 2 cout groups
    ('wg_pyr_up', 2, 256, 13, 21, 25, 42, 256, 2.5, False),
    ('wg_pyr_down', 2, 256, 25, 42, 13, 21, 256, 2.5, False),
    ('wg_level0ish', 1, 256, 50, 84, 50, 84, 256, 1.5, True),
]


@pytest.mark.parametrize('case', WG_CASES, ids=[c[0] for c in WG_CASES])
def test_fused_weight_gradient(case):
    """lsnet_dcn_backward_weight without saved columns: the fused kernel re-samples x into the MN-major B tiles.  Against
    (a) gather -> HBM columns -> split-K GEMM of the same library, (b) fp32 dY^T . columns; the deterministic two-stage
    reduction must be bit-reproducible."""
    ops, lib = _ops(), _lib()
    name, B, C, H, W, Ho, Wo, M, mag, has_mask = case
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    x = _nhwc(_bf(torch.randn(B, C, H, W, generator=g)), torch.bfloat16)
    off = _nhwc(torch.randn(B, 18, Ho, Wo, generator=g) * mag)
    mask = _nhwc(torch.rand(B, 9, Ho, Wo, generator=g)) if has_mask else None
    gy2 = _bf(torch.randn(B * Ho * Wo, M, generator=g)).to(DEV, torch.bfloat16)
    pyr = 'pyr' in name
    cfg = (Ho, Wo, 3, 3, (1, 1), (1, 1), (1, 1), (H / Ho, W / Wo) if pyr else (1.0, 1.0), 1)
    try:
        lib.lsnet_dcn_fused_enable(1)
        lib.lsnet_set_deterministic(0)
        dw_f = ops.dcn_backward_weight(gy2, x, off, mask, None, *cfg)
        acc = torch.full_like(dw_f, 0.5)
        ops.dcn_backward_weight(gy2, x, off, mask, None, *cfg, out=acc)          # accumulates into `out`
        lib.lsnet_set_deterministic(1)
        dw_d1 = ops.dcn_backward_weight(gy2, x, off, mask, None, *cfg)
        dw_d2 = ops.dcn_backward_weight(gy2, x, off, mask, None, *cfg)
        lib.lsnet_set_deterministic(0)
        lib.lsnet_dcn_fused_enable(0)
        col = ops.dcn_im2col(x, off, mask, *cfg)
        dw_u = ops.dcn_backward_weight(gy2, x, off, mask, col, *cfg)
        dw_u2 = ops.dcn_backward_weight(gy2, x, off, mask, None, *cfg)            # re-sampled through the workspace
    finally:
        lib.lsnet_dcn_fused_enable(2)
        lib.lsnet_set_deterministic(0)
    torch.cuda.synchronize()
    ref = gy2.float().t() @ col.float()
    assert torch.equal(dw_d1, dw_d2), name
    for nm, t in (('fused', dw_f), ('deterministic', dw_d1), ('columns', dw_u), ('resampled', dw_u2)):
        assert _rel(t, ref) < 2e-3, (name, nm, _rel(t, ref))
    assert _rel(acc - 0.5, ref) < 2e-3, name


@pytest.mark.parametrize('C,groups,stride', [(256, 64, 1), (512, 64, 2), (256, 8, 1)])
def test_grouped_dcn_native_vs_oracle(C, groups, stride):
    """DCNv2 with grouped weights on the library's block-diagonal kernels (the X-101-64x4d backbone sites, groups = 64,
    stride 2 in the first block of a stage: resnext.py:50-74, 114-118): forward, dX, dOffset, dMask, dW against the grouped
    oracle, and against the dense kernels on the zero-expanded weight."""
    ops = _ops()
    import lsnet_b200.ops.dcn as dcn_mod
    assert dcn_mod.grouped_native(C, C, 3, 3, groups, 1)
    B, H, W = 2, 12, 14
    Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    g = torch.Generator().manual_seed(C + groups + stride)
    x = _bf(torch.randn(B, C, H, W, generator=g))
    off = torch.randn(B, 18, Ho, Wo, generator=g) * 1.5
    mask = torch.rand(B, 9, Ho, Wo, generator=g)
    w = _bf(torch.randn(C, C // groups, 3, 3, generator=g) / (C // groups * 9) ** 0.5)
    gy = _bf(torch.randn(B, C, Ho, Wo, generator=g))
    ins_r = [t.clone().requires_grad_(True) for t in (x, off, mask, w)]
    ref = OD.modulated_deform_conv(*ins_r, None, stride, 1, 1, groups)
    rg = torch.autograd.grad(ref, ins_r, gy)
    ins = [t.to(DEV).requires_grad_(True) for t in (x, off, mask, w)]
    out = ops.modulated_deform_conv(*ins, None, stride, 1, 1, groups, out_fp32=True)
    assert out.shape == ref.shape
    assert _rel(out, ref) < 4e-3, _rel(out, ref)
    gg = torch.autograd.grad(out, ins, gy.to(DEV))
    for name, a, r in zip(['x', 'offset', 'mask', 'w'], gg, rg):
        assert a.shape == r.shape, name
        assert _rel(a.float(), r) < (4e-2 if name == 'x' else 2e-2), (C, groups, name, _rel(a.float(), r))
    # same op through the dense kernels on the expanded weight
    saved = dcn_mod.grouped_native
    dcn_mod.grouped_native = lambda *a, **k: False
    try:
        ins2 = [t.to(DEV).requires_grad_(True) for t in (x, off, mask, w)]
        out2 = ops.modulated_deform_conv(*ins2, None, stride, 1, 1, groups, out_fp32=True)
        gg2 = torch.autograd.grad(out2, ins2, gy.to(DEV))
    finally:
        dcn_mod.grouped_native = saved
    assert _rel(out, out2) < 2e-3
    assert _rel(gg[3], gg2[3]) < 5e-3


def test_dcn_pack_gradient_sink_matches_autograd_sum():
    """ModulatedDeformConvPack: x feeds conv_offset AND the sampling op.  With the gradient sink the conv_offset backward
    adds its input gradient into the sampling op's dX inside its GEMM epilogue (no separate sum); result must equal the
    plain autograd sum of the two gradients up to one bf16 rounding of the partial sum."""
    import torch
    from lsnet_b200.modules import dcn as D
    torch.manual_seed(3)
    m = D.ModulatedDeformConvPack(256, 256, 3, 1, 1).cuda()
    m.conv_offset.weight.data.normal_(0, 0.02)
    m.conv_offset.bias.data.normal_(0, 0.1)
    x0 = torch.randn(2, 256, 25, 42, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(2, 256, 25, 42, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    res = {}
    for mode in (True, False):
        D.GX_SINK = mode
        try:
            m.zero_grad()
            x = x0.clone().requires_grad_(True)
            m(x, exclusive=True).backward(gy)
            torch.cuda.synchronize()
            res[mode] = (x.grad.float().clone(), m.conv_offset.weight.grad.clone(), m.weight.grad.clone())
        finally:
            D.GX_SINK = True
    scale = res[False][0].abs().max()
    assert float((res[True][0] - res[False][0]).abs().max() / scale) < 1e-2
    assert torch.allclose(res[True][1], res[False][1], rtol=1e-3, atol=1e-4 * float(res[False][1].abs().max()))
    assert torch.allclose(res[True][2], res[False][2], rtol=1e-3, atol=1e-4 * float(res[False][2].abs().max()))


def test_dcn_pack_default_is_safe_with_outside_consumers():
    """Without the caller's ``exclusive`` promise the pack must not sum in place: x here also feeds an identity branch, so
    autograd accumulates an outside gradient with the sampling op's dX before conv_offset's arrives."""
    import torch
    from lsnet_b200.modules import dcn as D
    torch.manual_seed(5)
    m = D.ModulatedDeformConvPack(64, 64, 3, 1, 1).cuda()
    m.conv_offset.weight.data.normal_(0, 0.02)
    x0 = torch.randn(2, 64, 13, 21, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(2, 64, 13, 21, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    res = {}
    for mode in (True, False):
        D.GX_SINK = mode
        try:
            x = x0.clone().requires_grad_(True)
            (m(x) + x * 0.5).backward(gy)
            torch.cuda.synchronize()
            res[mode] = x.grad.float().clone()
        finally:
            D.GX_SINK = True
    assert float((res[True] - res[False]).abs().max() / res[False].abs().max()) < 1e-2


import pytest as _pytest


@_pytest.mark.parametrize('B,H,W', [(2, 13, 21), (4, 100, 168), (3, 7, 11)])
def test_gn_statistics_from_the_dcn_epilogue(B, H, W):
    """DCNConvModule = DCNv2 -> GroupNorm(32) -> ReLU (lsnet_head.py:1830-1849) with the GroupNorm statistics
    accumulated by the deformable convolution's GEMM epilogue (fused kernel on the large map, gather -> GEMM on the small
    ones, where a warp's 32 rows straddle image boundaries: 273 and 77 pixels per image) against the standalone statistics
    pass.  The epilogue sums the fp32 accumulator values, the standalone pass their bf16 roundings: outputs agree to
    bf16 rounding, gradients likewise."""
    import torch
    from lsnet_b200.modules import head as Hd
    torch.manual_seed(11)
    m = Hd.DCNConvModule(256, 256, 3, 1, 32, 1).cuda()
    m.conv.conv_offset.weight.data.normal_(0, 0.02)
    m.conv.bias.data.normal_(0, 0.5)
    m.bn.weight.data.uniform_(0.5, 1.5)
    m.bn.bias.data.normal_(0, 0.2)
    x0 = torch.randn(B, 256, H, W, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(B, 256, H, W, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    res = {}
    for mode in (True, False):
        Hd.GN_EPILOGUE = mode
        try:
            m.zero_grad()
            x = x0.clone().requires_grad_(True)
            y = m(x)
            y.backward(gy)
            torch.cuda.synchronize()
            res[mode] = (y.float().detach().clone(), x.grad.float().clone(), m.bn.weight.grad.clone(), m.conv.weight.grad.clone())
        finally:
            Hd.GN_EPILOGUE = True
    for name, a, b in zip(['y', 'dx', 'dgamma', 'dw'], res[True], res[False]):
        # dx is accumulated with bf16 reds whose order varies from run to run (two runs of the SAME mode differ by a few
        # per cent of max|dx| in single elements of a 27 M element map): compare it in the L2 norm
        # (the weight gradient of a 7x11 map sums 231 pixels: the handful of outputs whose ReLU mask flips with the last
        # bits of the statistics moves single entries by a few per cent -- gradients are compared in the L2 norm)
        err = float((a - b).abs().max() / (b.abs().max() + 1e-30)) if name == 'y' else float((a - b).norm() / (b.norm() + 1e-30))
        assert err < 2e-2, (name, err)
    # the statistics themselves: mean / variance of the normalised pre-activation are (beta, gamma^2) per group only in
    # expectation, so check against torch GroupNorm on the same DCN output instead
    Hd.GN_EPILOGUE = True
    h = {}
    raw = m.conv(x0, gn_holder=h, gn_groups=32)
    ref = torch.nn.functional.group_norm(raw.float(), 32, m.bn.weight, m.bn.bias, m.bn.eps).relu()
    y = m(x0)
    assert float((y.float() - ref).abs().max() / ref.abs().max()) < 1.5e-2
    sums = h['sums'][:2 * B * 32].view(B, 32, 2).float().cpu()
    r = raw.float().view(B, 32, 8, H * W)
    # tight enough to see ONE missing pixel row of a group (8 of 8 * H*W values: 1.3 % of the sum of squares at 7x11);
    # bf16 rounding of the reference values moves it by ~1e-4
    ss_ref = (r * r).sum((2, 3)).cpu()
    dev = float(((sums[..., 1] - ss_ref).abs() / ss_ref).max())
    assert dev < 2e-3, dev
    s_ref, s_abs = r.sum((2, 3)).cpu(), r.abs().sum((2, 3)).cpu()
    dev = float(((sums[..., 0] - s_ref).abs() / s_abs).max())
    assert dev < 2e-3, dev


def test_dcn_bias_gradient_from_the_norm_backward():
    """DCNConvModule under a trainer that exposes the parameters' gradient memory (GraphTrainer: ops.gemm_ops.direct_vec): the
    bias gradient of the deformable convolution is the per-channel sum of the GroupNorm's dx and is added by the norm's
    backward apply kernel; the convolution's own column-sum pass is skipped.  Same numbers as the autograd path."""
    import torch
    from lsnet_b200.modules import head as Hd
    torch.manual_seed(12)
    m = Hd.DCNConvModule(256, 256, 3, 1, 32, 1).cuda()
    m.conv.conv_offset.weight.data.normal_(0, 0.02)
    m.conv.bias.data.normal_(0, 0.5)
    m.bn.weight.data.uniform_(0.5, 1.5)
    x0 = torch.randn(2, 256, 25, 42, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(2, 256, 25, 42, device='cuda').to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    # reference: plain autograd accumulation
    m.zero_grad()
    m(x0.clone().requires_grad_(True)).backward(gy)
    ref = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    # direct mode: gradients of the 1-D parameters live in caller-owned memory the kernels add into
    m.zero_grad()
    for p in (m.bn.weight, m.bn.bias, m.conv.bias, m.conv.conv_offset.bias):
        p.grad = torch.zeros_like(p)
        p._lsnet_direct_vec = True
    from lsnet_b200 import ops
    assert ops.norm.bias_sink_ok(m.bn.weight, m.bn.bias, m.conv.bias)
    m(x0.clone().requires_grad_(True)).backward(gy)
    torch.cuda.synchronize()
    for k in ('conv.bias', 'bn.weight', 'bn.bias', 'conv.conv_offset.bias', 'conv.weight'):
        got, want = dict(m.named_parameters())[k].grad, ref[k]
        assert float((got - want).abs().max() / (want.abs().max() + 1e-30)) < 2e-2, k
    for p in (m.bn.weight, m.bn.bias, m.conv.bias, m.conv.conv_offset.bias):
        p._lsnet_direct_vec = False
