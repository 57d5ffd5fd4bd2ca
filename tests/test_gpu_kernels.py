"""GPU parity tests (run on the B200 box): every kernel of liblsnet_sm100.so, called through the C ABI, against the
oracle (oracle/ = CPU restatement of the reference) on identical seeded inputs.

Tolerances.  Integer / index outputs (assignment, labels, slot masks): bit-exact.  fp32 loss kernels: 1e-4 relative
(north_star).  bf16 tensor-core paths (GEMM / conv / DCN): operands are rounded to bf16 BEFORE both sides compute,
accumulation is fp32 on both sides, so with fp32 outputs the only differences are summation order (<= 2e-3 of the
output scale, K up to 2304); the DCN column values are additionally rounded to bf16 (2^-9 relative per element),
bounded the same way; bf16 outputs add one more 2^-9 rounding (bound 1.5e-2).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import synth
from oracle import dcn_ops as OD
from oracle import lsnet_oracle as O

pytestmark = pytest.mark.gpu

DEV = 'cuda'


def _ops():
    import lsnet_b200.ops as ops
    return ops


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _bf(t):
    return t.to(torch.bfloat16).float()


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize('M,N,K', [(128, 256, 64), (300, 256, 256), (1000, 32, 2304), (257, 64, 768), (4200, 128, 512),
                                   (16800, 256, 2304), (2000, 2304, 256), (77, 80, 256), (5, 16, 8)])
def test_gemm_kmajor(M, N, K):
    ops = _ops()
    g = torch.Generator().manual_seed(M + N + K)
    a = _bf(torch.randn(M, K, generator=g))
    b = _bf(torch.randn(N, K, generator=g))
    bias = torch.randn(N, generator=g)
    ref = a @ b.t() + bias
    out = ops.gemm(a.to(DEV, torch.bfloat16), b.to(DEV, torch.bfloat16), bias.to(DEV), False, torch.float32)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 2e-3
    out = ops.gemm(a.to(DEV, torch.bfloat16), b.to(DEV, torch.bfloat16), bias.to(DEV), True, torch.bfloat16)
    assert _rel(out.float(), ref.clamp(min=0)) < 1.5e-2


@pytest.mark.parametrize('P,M,N', [(64, 128, 256), (1000, 256, 2304), (16800, 256, 256), (273, 32, 256), (70, 8, 24)])
def test_gemm_tn(P, M, N):
    ops = _ops()
    g = torch.Generator().manual_seed(P + M + N)
    a = _bf(torch.randn(P, M, generator=g))
    b = _bf(torch.randn(P, N, generator=g))
    ref = a.t() @ b
    out = ops.gemm_tn(a.to(DEV, torch.bfloat16), b.to(DEV, torch.bfloat16))
    assert _rel(out, ref) < 2e-3


@pytest.mark.parametrize('B,C,H,W,N,k', [(2, 256, 13, 21, 256, 3), (1, 256, 25, 42, 32, 3), (2, 64, 7, 11, 256, 3),
                                         (1, 32, 9, 9, 256, 3), (1, 256, 50, 84, 256, 3), (2, 512, 10, 12, 256, 1),
                                         (1, 768, 13, 21, 256, 1)])
def test_conv2d_same_forward_backward(B, C, H, W, N, k):
    ops = _ops()
    g = torch.Generator().manual_seed(B * C + H * W + N)
    x = _bf(torch.randn(B, C, H, W, generator=g)).requires_grad_(True)
    w = _bf(torch.randn(N, C, k, k, generator=g) / (C * k * k) ** 0.5).requires_grad_(True)
    bias = torch.randn(N, generator=g).requires_grad_(True)
    gy = _bf(torch.randn(B, N, H, W, generator=g))
    ref = F.conv2d(x, w, bias, 1, k // 2)
    rgx, rgw, rgb = torch.autograd.grad(ref, [x, w, bias], gy)
    xd = x.detach().to(DEV).requires_grad_(True)
    wd = w.detach().to(DEV).requires_grad_(True)
    bd = bias.detach().to(DEV).requires_grad_(True)
    out = ops.conv2d_same(xd, wd, bd, padding=k // 2, out_fp32=True)
    assert out.shape == ref.shape
    assert _rel(out, ref) < 2e-3
    gx, gw, gb = torch.autograd.grad(out, [xd, wd, bd], gy.to(DEV))
    assert _rel(gx.float(), rgx) < 1.5e-2       # bf16 dX
    assert _rel(gw, rgw) < 2e-3
    assert _rel(gb, rgb) < 1e-4


# ------------------------------------------------------------------------------------------------ DCN
def _dcn_case(seed, B, C, H, W, Ho, Wo, Co, mag):
    g = torch.Generator().manual_seed(seed)
    x = _bf(torch.randn(B, C, H, W, generator=g))
    off = torch.randn(B, 18, Ho, Wo, generator=g) * mag
    mask = torch.rand(B, 9, Ho, Wo, generator=g)
    w = _bf(torch.randn(Co, C, 3, 3, generator=g) / (C * 9) ** 0.5)
    bias = torch.randn(Co, generator=g)
    gy = _bf(torch.randn(B, Co, Ho, Wo, generator=g))
    return x, off, mask, w, bias, gy


@pytest.mark.parametrize('dx_fp32', [True, False])
@pytest.mark.parametrize('variant', ['v2', 'v1', 'pyramid_up', 'pyramid_down', 'pyramid_same'])
def test_dcn_forward_backward_vs_oracle(variant, dx_fp32):
    """dX accumulation: fp32 reds (2e-2 of the gradient scale, dominated by the bf16 dCol operand) or the training
    default packed-bf16 reds, whose running sum is rounded to bf16 at every add: up to ~sqrt(n) * 2^-9 for n
    contributions per input pixel (n ~ 36 at scale 1, ~144 when the offset grid is 2x finer than the input) -> 4e-2."""
    ops = _ops()
    import lsnet_b200.ops.dcn as dcn_mod
    dcn_mod.DX_FP32 = dx_fp32
    B, C, Co = 2, 64, 48
    if variant in ('v2', 'v1'):
        H, W, Ho, Wo = 13, 21, 13, 21
    elif variant == 'pyramid_up':      # output grid finer than the sampled map (non-dyadic 13->25, 21->42)
        H, W, Ho, Wo = 13, 21, 25, 42
    elif variant == 'pyramid_down':
        H, W, Ho, Wo = 25, 42, 13, 21
    else:
        H, W, Ho, Wo = 9, 10, 9, 10
    x, off, mask, w, bias, gy = _dcn_case(hash(variant) % 1000, B, C, H, W, Ho, Wo, Co, 2.5)
    sc = (H / Ho, W / Wo)
    xr, offr, maskr, wr, br = (t.clone().requires_grad_(True) for t in (x, off, mask, w, bias))
    if variant == 'v2':
        ref = OD.modulated_deform_conv(xr, offr, maskr, wr, br, 1, 1, 1)
        ins_r = [xr, offr, maskr, wr, br]
    elif variant == 'v1':
        ref = OD.deform_conv(xr, offr, wr, 1, 1, 1)
        ins_r = [xr, offr, wr]
    else:
        ref = OD.pyramid_deform_conv(xr, offr, wr, sc, 1, 1, 1)
        ins_r = [xr, offr, wr]
    rg = torch.autograd.grad(ref, ins_r, gy)
    xd, offd, maskd, wd, bd = (t.to(DEV).requires_grad_(True) for t in (x, off, mask, w, bias))
    if variant == 'v2':
        out = ops.modulated_deform_conv(xd, offd, maskd, wd, bd, 1, 1, 1, out_fp32=True)
        ins = [xd, offd, maskd, wd, bd]
    elif variant == 'v1':
        out = ops.deform_conv(xd, offd, wd, 1, 1, 1, out_fp32=True)
        ins = [xd, offd, wd]
    else:
        out = ops.pyramid_deform_conv(xd, offd, wd, sc, 1, 1, 1, out_fp32=True)
        ins = [xd, offd, wd]
    assert out.shape == ref.shape
    assert _rel(out, ref) < 4e-3
    gg = torch.autograd.grad(out, ins, gy.to(DEV))
    for name, a, r in zip(['x', 'offset', 'mask/w', 'w/b', 'b'], gg, rg):
        assert a.shape == r.shape, name
        tol = 4e-2 if (name == 'x' and not dx_fp32) else 2e-2
        assert _rel(a.float(), r) < tol, (variant, name, _rel(a.float(), r))
    dcn_mod.DX_FP32 = False


@pytest.mark.parametrize('groups,stride', [(4, 1), (8, 2)])
def test_dcn_grouped_vs_oracle(groups, stride):
    """DCNv2 with conv groups > 1 (the X-101-64x4d backbone sites: groups=64, stride 2 in the first block of a stage,
    deform_conv.py:488-534 called from resnext.py:114-118) against the grouped oracle."""
    ops = _ops()
    B, C, Co, H, W = 2, 64, 64, 12, 14
    Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    g = torch.Generator().manual_seed(77 + groups)
    x = _bf(torch.randn(B, C, H, W, generator=g))
    off = torch.randn(B, 18, Ho, Wo, generator=g) * 1.5
    mask = torch.rand(B, 9, Ho, Wo, generator=g)
    w = _bf(torch.randn(Co, C // groups, 3, 3, generator=g) / (C // groups * 9) ** 0.5)
    gy = _bf(torch.randn(B, Co, Ho, Wo, generator=g))
    ins_r = [t.clone().requires_grad_(True) for t in (x, off, mask, w)]
    ref = OD.modulated_deform_conv(*ins_r, None, stride, 1, 1, groups)
    rg = torch.autograd.grad(ref, ins_r, gy)
    ins = [t.to(DEV).requires_grad_(True) for t in (x, off, mask, w)]
    out = ops.modulated_deform_conv(*ins, None, stride, 1, 1, groups, out_fp32=True)
    assert out.shape == ref.shape
    assert _rel(out, ref) < 4e-3
    gg = torch.autograd.grad(out, ins, gy.to(DEV))
    for name, a, r in zip(['x', 'offset', 'mask', 'w'], gg, rg):
        assert a.shape == r.shape, name
        assert _rel(a.float(), r) < (4e-2 if name == 'x' else 2e-2), (groups, name, _rel(a.float(), r))


def test_dcnv2_pack_fused_sigmoid_matches_unfused():
    """ModulatedDeformConvPack (deform_conv.py:488-534): the one-op path (sampling kernels read offsets and mask LOGITS
    straight from the conv_offset output, sigmoid and its derivative inside) against chunk/cat/sigmoid in torch."""
    import lsnet_b200.modules.dcn as mdcn
    torch.manual_seed(4)
    m = mdcn.ModulatedDeformConvPack(64, 48, 3, stride=1, padding=1, bias=True).to(DEV)
    torch.nn.init.normal_(m.conv_offset.weight, std=0.05)
    torch.nn.init.normal_(m.conv_offset.bias, std=0.5)
    x = _bf(torch.randn(2, 64, 13, 21)).to(DEV)
    gy = _bf(torch.randn(2, 48, 13, 21)).to(DEV)
    res = []
    for packed in (True, False):
        mdcn.PACKED_OFFSET_MASK = packed
        xi = x.clone().requires_grad_(True)
        m.zero_grad()
        y = m(xi)
        y.backward(gy.to(y.dtype))
        res.append((y.detach().float(), xi.grad.float(), m.conv_offset.weight.grad.clone(), m.conv_offset.bias.grad.clone(),
                    m.weight.grad.clone()))
    mdcn.PACKED_OFFSET_MASK = True
    for name, a, b in zip(['y', 'dx', 'd conv_offset.weight', 'd conv_offset.bias', 'dw'], res[0], res[1]):
        assert _rel(a, b) < 2e-2, (name, _rel(a, b))      # bf16 dX reds / column rounding order
    assert _rel(res[0][0], res[1][0]) < 1e-3 and _rel(res[0][3], res[1][3]) < 1e-3


@pytest.mark.parametrize('variant', ['v2', 'pyramid_half', 'pyramid_double'])
def test_dcn_adjoint_identity_full_size(variant):
    """BASELINE cfg2 sizes (B=4, 256 ch, level-0 / level-1 grids): the scatter is the exact adjoint of the gather.  With
    J = d(columns)/d(x) for fixed offsets, <J^T g, x'> == <g, J x'> for arbitrary g, x' — a size-independent property
    that needs no oracle.  Checked in fp64 on the fp32-accumulated dX; the only inexact step is the bf16 rounding of
    the gathered columns (2^-9 relative per element, random sign)."""
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    B, C = 4, 256
    if variant == 'v2':
        H, W, Ho, Wo, sc = 100, 168, 100, 168, (1.0, 1.0)
    elif variant == 'pyramid_half':      # level-0 grid sampling level 1
        H, W, Ho, Wo, sc = 50, 84, 100, 168, (0.5, 0.5)
    else:                                # level-1 grid sampling level 0
        H, W, Ho, Wo, sc = 100, 168, 50, 84, (2.0, 2.0)
    xp = torch.randn(B, C, H, W, generator=g).to(DEV, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    off = (torch.randn(B, 18, Ho, Wo, generator=g) * 1.5).to(DEV).contiguous(memory_format=torch.channels_last)
    mask = torch.rand(B, 9, Ho, Wo, generator=g).to(DEV).contiguous(memory_format=torch.channels_last) if variant == 'v2' else None
    gcol = torch.randn(B * Ho * Wo, 9 * C, generator=g).to(DEV, torch.bfloat16)
    cfg = (Ho, Wo, 3, 3, (1, 1), (1, 1), (1, 1), sc, 1)
    col = ops.dcn_im2col(xp, off, mask, *cfg)                       # J x'  (bf16-rounded)
    dx, doff, dmask = ops.dcn_col2im(gcol, xp, off, mask, *cfg, dx_fp32=True)       # J^T g
    lhs = float((dx.double() * xp.double()).sum())
    rhs = float((gcol.double() * col.double()).sum())
    scale = float(gcol.double().norm() * col.double().norm())
    # rounding noise of the columns: independent 2^-9 relative errors -> std = 2^-9 * |g||col| / sqrt(n); allow 8 sigma
    tol = 8 * 2.0 ** -9 * scale / gcol.numel() ** 0.5
    assert abs(lhs - rhs) < tol, (variant, lhs, rhs, tol)
    assert torch.isfinite(doff).all() and (dmask is None or torch.isfinite(dmask).all())


def test_dcn_out_of_range_and_zero_offsets():
    """P4 / P6: far-away samples contribute nothing; zero offsets + mask 0.5 reduce DCNv2 to 0.5 * conv."""
    ops = _ops()
    x, off, mask, w, bias, gy = _dcn_case(5, 1, 64, 8, 9, 8, 9, 32, 0.0)
    out = ops.modulated_deform_conv(x.to(DEV), torch.zeros_like(off).to(DEV), torch.full_like(mask, 0.5).to(DEV),
                                    w.to(DEV), None, 1, 1, 1, out_fp32=True)
    assert _rel(out, 0.5 * F.conv2d(x, w, None, 1, 1)) < 4e-3
    out = ops.deform_conv(x.to(DEV), torch.full_like(off, 100.0).to(DEV), w.to(DEV), 1, 1, 1, out_fp32=True)
    assert float(out.abs().max()) == 0.0


def test_dcn_full_size_properties():
    """BASELINE cfg2 level-0 shape (B=4, 256 ch, 100x168): size-independent properties.  (i) Scaling x or W by 2 is
    exact in bf16 and fp32, so the op must commute with it BIT-exactly; (ii) zero offsets + unit mask reduce DCNv2 to
    the plain 3x3 convolution (checked against cuDNN on the same bf16 operands)."""
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    B, C, H, W = 4, 256, 100, 168
    x = _bf(torch.randn(B, C, H, W, generator=g)).to(DEV)
    off = (torch.randn(B, 18, H, W, generator=g) * 1.5).to(DEV)
    mask = torch.rand(B, 9, H, W, generator=g).to(DEV)
    w = _bf(torch.randn(256, 256, 3, 3, generator=g) / 48).to(DEV)
    y = ops.modulated_deform_conv(x, off, mask, w, None, 1, 1, 1, out_fp32=True)
    assert torch.equal(ops.modulated_deform_conv(2 * x, off, mask, w, None, 1, 1, 1, out_fp32=True), 2 * y)
    assert torch.equal(ops.modulated_deform_conv(x, off, mask, 2 * w, None, 1, 1, 1, out_fp32=True), 2 * y)
    y0 = ops.modulated_deform_conv(x, torch.zeros_like(off), torch.ones_like(mask), w, None, 1, 1, 1, out_fp32=True)
    ref = F.conv2d(x, w, None, 1, 1)
    assert _rel(y0, ref) < 4e-3
    yc = ops.conv2d_same(x, w, None, padding=1, out_fp32=True)
    assert _rel(yc, ref) < 2e-3


# ------------------------------------------------------------------------------------------------ losses
@pytest.mark.parametrize('lt', ['bbox', 'polygon', 'keypoint'])
def test_cross_iou_dense_100_steps(lt):
    """Loss and gradient within 1e-4 relative of the oracle over 100 synthetic steps (north_star)."""
    ops = _ops()
    for step in range(100):
        r = synth.loss_rows(lt, 3000 + step, N=48)
        tgt, sel = O.directional_targets(r['gt'], r['anchor'], r['weight'])
        pred = r['pred'].clone().requires_grad_(True)
        kw = dict(loss_type=lt, anchor_pts=r['anchor'], bbox_gt=None if lt == 'keypoint' else r['bbox_gt'],
                  pos_inds=sel)
        if lt == 'keypoint':
            kw['vs'] = r['vs']
        ref = O.cross_iou_loss(pred, tgt, r['weight'], 11.0, 2.0, **kw)
        ref.backward()
        pd = r['pred'].to(DEV).requires_grad_(True)
        rows = ops.cross_iou_loss_rows(pd, tgt.to(DEV), sel.to(DEV), r['weight'].mean(1).to(DEV), r['anchor'].to(DEV),
                                       None if lt == 'keypoint' else r['bbox_gt'].to(DEV),
                                       r['vs'].to(DEV) if lt == 'keypoint' else None, loss_type=lt)
        loss = 2.0 * rows.sum() / 11.0
        loss.backward()
        assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref)), (step, float(loss), float(ref))
        gn = (pd.grad.cpu() - pred.grad).norm() / pred.grad.norm()
        assert float(gn) <= 1e-4, (step, float(gn))


def test_directional_targets_bit_exact():
    ops = _ops()
    for lt in ('bbox', 'polygon'):
        r = synth.loss_rows(lt, 777)
        t, s = O.directional_targets(r['gt'], r['anchor'], r['weight'])
        td, sd = ops.directional_targets(r['gt'].to(DEV), r['anchor'].to(DEV), r['weight'][:, 0].to(DEV))
        assert torch.equal(td.cpu(), t) and torch.equal(sd.cpu(), s)


def test_focal_vs_oracle():
    ops = _ops()
    rng = np.random.RandomState(9)
    logits = torch.from_numpy((rng.randn(5000, 80) * 3).astype(np.float32))
    labels = torch.from_numpy(rng.randint(0, 81, 5000))
    w = torch.from_numpy((rng.rand(5000) > 0.1).astype(np.float32))
    lr = logits.clone().requires_grad_(True)
    ref = O.focal_loss(lr, labels, w, 13.0)
    ref.backward()
    ld = logits.to(DEV).requires_grad_(True)
    loss = ops.sigmoid_focal_loss_sum(ld, labels.to(DEV), w.to(DEV)) / 13.0
    loss.backward()
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref))
    assert float((ld.grad.cpu() - lr.grad).norm() / lr.grad.norm()) <= 1e-4


# ------------------------------------------------------------------------------------------------ assignment
def _gt_pack(gts, dev):
    B, Gmax = len(gts), max(1, max(len(g) for g in gts))
    bb = torch.zeros(B, Gmax, 4)
    cnt = torch.zeros(B, dtype=torch.int32)
    for i, g in enumerate(gts):
        bb[i, :len(g)] = g
        cnt[i] = len(g)
    return bb.to(dev), cnt.to(dev)


def test_assignment_bit_exact_vs_oracle_and_golden(golden_dir):
    import os
    ops = _ops()
    gold = np.load(os.path.join(golden_dir, 'assign.npz'))
    cases = [synth.assign_case(900 + s, G=1 + 3 * s) for s in range(6)]
    c0 = cases[0]
    pyr = ops.Pyramid(c0['sizes'], c0['strides'], [(384, 448)] * len(cases), DEV)
    bb, cnt = _gt_pack([c['gt'] for c in cases], DEV)
    a1 = ops.centroid_assign(pyr, bb, cnt).cpu()
    boxes = torch.stack([c['pred'] for c in cases]).to(DEV)
    a2, mo = ops.atss_assign(pyr, boxes, bb, cnt, want_overlaps=True)
    a2, mo = a2.cpu(), mo.cpu()
    for s, c in enumerate(cases):
        r1 = O.centroid_assign(c['points'], c['gt'])
        r2, rmo = O.atss_assign(c['pred'], c['num_level'], c['gt'])
        assert np.array_equal(r1.numpy(), gold[f'{s}.centroid']) and np.array_equal(r2.numpy(), gold[f'{s}.atss'])
        assert torch.equal(a1[s].long() + 1, r1), f'centroid case {s}'
        assert torch.equal(a2[s].long() + 1, r2), f'atss case {s}'
        pos = r2 > 0
        assert torch.equal(mo[s][pos], rmo[pos])


def test_assignment_invalid_points_and_empty_gt():
    """P13: points outside an image's pad_shape are never assigned; an image without GT is all background."""
    ops = _ops()
    c = synth.assign_case(42, G=5)
    pad = (300, 330)
    pyr = ops.Pyramid(c['sizes'], c['strides'], [pad, (384, 448)], DEV)
    bb, cnt = _gt_pack([c['gt'], torch.zeros(0, 4)], DEV)
    valid = torch.cat([O.valid_flags(h, w, min(int(np.ceil(pad[0] / s)), h), min(int(np.ceil(pad[1] / s)), w))
                       for (h, w), s in zip(c['sizes'], c['strides'])])
    a1 = ops.centroid_assign(pyr, bb, cnt).cpu()
    ref = torch.zeros(len(valid), dtype=torch.long)
    ref[valid] = O.centroid_assign(c['points'][valid], c['gt'])
    assert torch.equal(a1[0].long() + 1, ref)
    assert int((a1[1] >= 0).sum()) == 0
    boxes = torch.stack([c['pred'], c['pred']]).to(DEV)
    a2 = ops.atss_assign(pyr, boxes, bb, cnt).cpu()
    inside = [int(f.sum()) for f in torch.split(valid, c['num_level'])]
    ref2 = torch.zeros(len(valid), dtype=torch.long)
    ref2[valid] = O.atss_assign(c['pred'][valid], inside, c['gt'])[0]
    assert torch.equal(a2[0].long() + 1, ref2)
    assert int((a2[1] >= 0).sum()) == 0
    labels, lw, npos = ops.assign_targets(pyr, a2.to(DEV), torch.arange(5, dtype=torch.int32).repeat(2, 1).to(DEV), 80)
    assert int(npos[0]) == int((ref2 > 0).sum()) and int(npos[1]) == 0
    assert torch.equal(lw[0].cpu() > 0, valid)
    lab_ref = torch.full((len(valid),), 80, dtype=torch.long)
    lab_ref[ref2 > 0] = ref2[ref2 > 0] - 1
    assert torch.equal(labels[0].cpu().long(), lab_ref)


def test_pred_boxes_vs_oracle():
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    sizes, strides = [(12, 14), (6, 7)], [8, 16]
    preds = [F.softplus(torch.randn(2, 20, h, w, generator=g)) for h, w in sizes]
    pyr = ops.Pyramid(sizes, strides, [(96, 112)] * 2, DEV)
    pd = [p.to(DEV).contiguous(memory_format=torch.channels_last) for p in preds]
    boxes = ops.pred_boxes(pyr, pd).cpu()
    off = 0
    for p, (h, w), s in zip(preds, sizes, strides):
        pts = O.grid_points(h, w, s)
        ref = torch.cat([pts[:, :2], pts[:, :2]], 1)[None] + (O.extreme_points2bbox(p) * s).permute(0, 2, 3, 1).reshape(2, -1, 4)
        assert torch.equal(boxes[:, off:off + h * w], ref)
        off += h * w


# ------------------------------------------------------------------------------------------------ GroupNorm
@pytest.mark.parametrize('B,C,H,W,relu,res', [(2, 256, 13, 21, True, False), (1, 256, 50, 84, True, True),
                                               (3, 256, 7, 11, False, False), (2, 64, 9, 10, False, True)])
def test_groupnorm_forward_backward(B, C, H, W, relu, res):
    """GN(32 groups) (+residual, +ReLU) vs torch's fp32 group_norm on the same bf16-rounded inputs."""
    ops = _ops()
    g = torch.Generator().manual_seed(B * C + H)
    G = 32 if C >= 256 else 8
    x = _bf(torch.randn(B, C, H, W, generator=g) * 2 + 0.3)
    x2 = _bf(torch.randn(B, C, H, W, generator=g)) if res else None
    w = torch.randn(C, generator=g) * 0.5 + 1
    b = torch.randn(C, generator=g) * 0.2
    gy = _bf(torch.randn(B, C, H, W, generator=g))
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    x2r = x2.clone().requires_grad_(True) if res else None
    s = _bf(xr + x2r) if res else xr
    if res:   # keep autograd through the rounding as identity
        s = (xr + x2r) + (_bf(xr + x2r) - (xr + x2r)).detach()
    ref = F.group_norm(s, G, wr, br, 1e-5)
    ref = F.relu(ref) if relu else ref
    ins_r = [xr, wr, br] + ([x2r] if res else [])
    rg = torch.autograd.grad(ref, ins_r, gy)
    xd, wd, bd = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    x2d = x2.to(DEV).requires_grad_(True) if res else None
    out = ops.group_norm_nhwc(xd, G, wd, bd, 1e-5, relu=relu, residual=x2d)
    assert _rel(out.float(), ref) < 1.5e-2
    gg = torch.autograd.grad(out, [xd, wd, bd] + ([x2d] if res else []), gy.to(DEV))
    assert _rel(gg[0].float(), rg[0]) < 2e-2
    assert _rel(gg[1], rg[1]) < 1e-2 and _rel(gg[2], rg[2]) < 1e-2
    if res:
        assert _rel(gg[3].float(), rg[3]) < 2e-2


def test_fused_clip_sgd_step_matches_torch():
    """lsnet_sgd_momentum_step (clip at max_norm + SGD momentum/weight-decay in one pass over the flat buffers) against
    clip_grad_norm_ + torch.optim.SGD (mmcv OptimizerHook.after_train_iter, optimizer.py:19-28) over three steps."""
    from lsnet_b200 import lib as L
    torch.manual_seed(0)
    n = 64 * 1000
    p0 = torch.randn(n, device=DEV)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([ref_p], lr=0.01, momentum=0.9, weight_decay=1e-4)
    p, m = p0.clone(), torch.zeros(n, device=DEV)
    for step, gscale in enumerate([100.0, 0.001, 10.0]):      # clipped, unclipped, clipped
        g = torch.randn(n, device=DEV) * gscale
        ref_p.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([ref_p], 35.0)
        opt.step()
        norm = torch.linalg.vector_norm(g)
        L.call('lsnet_sgd_momentum_step', L.ptr(p), L.ptr(g), L.ptr(m), L.c_ll(n), L.ptr(norm), L.c_f(35.0), L.c_f(0.01),
               L.c_f(0.9), L.c_f(1e-4), L.stream())
        assert _rel(p, ref_p.detach()) < 1e-6, (step, _rel(p, ref_p.detach()))


@pytest.mark.parametrize('branch,num_vectors,n_out', [('bbox', 4, 28), ('segm', 36, 148), ('pose', 17, 72)])
def test_head_glue_kernels_vs_torch(branch, num_vectors, n_out):
    """lsnet_pred_reg_fwd/bwd and lsnet_add_softplus against the reference's torch formulation
    (LSHead.get_pred_reg + softplus + gradient-mul mix, lsnet_head.py:372-400,585-587,735-755), values and gradients."""
    import math
    ops = _ops()
    gmul = 0.1
    g = torch.Generator().manual_seed(5)
    o = (torch.randn(2, n_out, 7, 9, generator=g) * 3).to(DEV)
    o[0, 0, 0, 0] = 25.0                      # softplus threshold branch
    o[0, 2, 1, 1] = o[0, 3, 1, 1]             # tie inside a pair: the '-' slot wins
    base = [float(v) for v in np.stack([np.repeat(np.arange(-1, 2.), 3), np.tile(np.arange(-1, 2.), 3)], 1).reshape(-1)]
    n_sp, src, mode = ops.pred_reg_table(branch, num_vectors, 9, n_out)

    def signed_pairs(t):
        r = t.reshape(t.shape[0], -1, 2, *t.shape[2:])
        val, ind = r.max(dim=2)
        return torch.where(ind == 0, -val, val)

    def ref_fn(x):
        sp = F.softplus(x[:, :n_sp])
        if branch == 'bbox':
            reg = torch.cat((signed_pairs(sp), x[:, n_sp:]), 1)
        else:
            r = sp.reshape(sp.shape[0], -1, 4, *sp.shape[2:])
            cts, polys = r[:, -1:], r[:, :-1]
            sel = polys[:, ::math.ceil(num_vectors / 8)] if branch == 'segm' else polys[:, 1::2]
            offs = torch.cat([sel, cts], 1)
            reg = signed_pairs(offs.reshape(offs.shape[0], -1, *offs.shape[3:]))
        reg = (1 - gmul) * reg.detach() + gmul * reg
        return sp, reg - torch.tensor(base, device=x.device).view(1, -1, 1, 1)

    gsp = torch.randn(2, n_sp, 7, 9, generator=g).to(DEV)
    goff = torch.randn(2, 18, 7, 9, generator=g).to(DEV)
    xr = o.clone().requires_grad_(True)
    sp_r, off_r = ref_fn(xr)
    (sp_r * gsp).sum().add((off_r * goff).sum()).backward()
    xg = o.clone().requires_grad_(True)
    sp, off = ops.pred_reg(xg, n_sp, src, mode, base, gmul)
    assert sp.shape == sp_r.shape and off.shape == off_r.shape
    assert torch.allclose(sp, sp_r, rtol=1e-6, atol=1e-6) and torch.allclose(off, off_r, rtol=1e-6, atol=1e-6)
    (sp * gsp).sum().add((off * goff).sum()).backward()
    assert torch.allclose(xg.grad, xr.grad, rtol=1e-5, atol=1e-6), float((xg.grad - xr.grad).abs().max())
    # refine = softplus(raw + init.detach())
    t = torch.randn(2, n_sp, 7, 9, generator=g).to(DEV)
    tr, tg = t.clone().requires_grad_(True), t.clone().requires_grad_(True)
    yr = F.softplus(tr + sp_r.detach())
    yg = ops.add_softplus(tg, sp.detach())
    assert torch.allclose(yg, yr, rtol=1e-6, atol=1e-6)
    (yr * gsp).sum().backward()
    (yg * gsp).sum().backward()
    assert torch.allclose(tg.grad, tr.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('Hf,Wf,Hc,Wc', [(100, 168, 50, 84), (25, 42, 13, 21), (13, 21, 7, 11), (9, 9, 4, 5)])
def test_upsample_add_matches_interpolate(Hf, Wf, Hc, Wc):
    """FPN top-down step (necks/fpn.py:180-192): fine + F.interpolate(coarse, size=fine.shape, mode='nearest') and its
    gradients -- bit-equal forward (one fp32 add, one rounding), gradient of the coarse map to bf16 rounding of the sum."""
    ops = _ops()
    g = torch.Generator().manual_seed(Hf * Wf)
    fine = _bf(torch.randn(2, 64, Hf, Wf, generator=g)).to(DEV)
    coarse = _bf(torch.randn(2, 64, Hc, Wc, generator=g)).to(DEV)
    gy = _bf(torch.randn(2, 64, Hf, Wf, generator=g)).to(DEV)
    fr, cr = fine.clone().requires_grad_(True), coarse.clone().requires_grad_(True)
    ref = fr + F.interpolate(cr, size=(Hf, Wf), mode='nearest')
    ref.backward(gy)
    fo = fine.to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    co = coarse.to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    out = ops.upsample_add(fo, co)
    out.backward(gy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
    torch.cuda.synchronize()
    assert torch.equal(out.float(), _bf(ref.detach()))
    assert torch.equal(fo.grad.float(), gy)
    assert _rel(co.grad.float(), cr.grad) < 8e-3
