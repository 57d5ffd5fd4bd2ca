"""Host logic of the strided-convolution path (no GPU): the phase decomposition of the input gradient and the drop-in
deform_conv_ext signatures."""
import inspect
import itertools

import numpy as np
import pytest


@pytest.mark.parametrize('k,stride,pad,dil', [(3, 2, 1, 1), (1, 2, 0, 1), (3, 1, 1, 1), (3, 1, 2, 2), (5, 2, 2, 1), (3, 3, 1, 1),
                                              (7, 2, 3, 1)])
def test_dgrad_phases_reproduce_the_transposed_convolution(k, stride, pad, dil):
    """ops.conv.dgrad_phases: phase (ph, pw) of the input gradient reads dY at (i + dy, j + dx) with weight block t for
    every listed tap.  Checked against the definition gx[h] = sum_{ho, t : ho*s - p + t*d == h} gy[ho] * w[t] in 1-D x 1-D
    (the 2-D table is the outer product of the two 1-D ones) on random data."""
    from lsnet_b200.ops.conv import conv_out_hw, dgrad_phases
    rng = np.random.RandomState(k * 100 + stride * 10 + pad)
    H, W = 11, 9
    Ho, Wo = conv_out_hw(H, W, k, k, stride, pad, dil)
    gy = rng.randn(Ho, Wo)
    w = rng.randn(k, k)
    ref = np.zeros((H, W))
    for ho, wo, ty, tx in itertools.product(range(Ho), range(Wo), range(k), range(k)):
        h, x = ho * stride - pad + ty * dil, wo * stride - pad + tx * dil
        if 0 <= h < H and 0 <= x < W:
            ref[h, x] += gy[ho, wo] * w[ty, tx]
    got = np.zeros((H, W))
    for ph, pw, taps in dgrad_phases(k, k, stride, pad, dil):
        Hp, Wp = (H - ph + stride - 1) // stride, (W - pw + stride - 1) // stride
        for i, j in itertools.product(range(max(Hp, 0)), range(max(Wp, 0))):
            acc = 0.0
            for dy, dx, kb in taps:
                a, b = i + dy, j + dx
                if 0 <= a < Ho and 0 <= b < Wo:          # out-of-range taps are TMA zero fill
                    acc += gy[a, b] * w[kb // k, kb % k]
            got[i * stride + ph, j * stride + pw] = acc
    assert np.allclose(got, ref, atol=1e-12)
    # every tap belongs to exactly one phase
    seen = sorted(kb for _, _, taps in dgrad_phases(k, k, stride, pad, dil) for _, _, kb in taps)
    assert seen == list(range(k * k))


def test_compat_deform_conv_ext_signatures():
    """lsnet_b200.compat.deform_conv_ext exports the reference's eight pybind functions with the same positional argument
    counts (mmdet/ops/dcn/src/deform_conv_ext.cpp:74-224: 17 / 18 / 18 / 19 / 24 / 19 / 20 / 20 arguments)."""
    from lsnet_b200.compat import deform_conv_ext as m
    want = dict(deform_conv_forward=17, deform_conv_backward_input=18, deform_conv_backward_parameters=18,
                modulated_deform_conv_forward=19, modulated_deform_conv_backward=24, pyramid_deform_conv_forward=19,
                pyramid_deform_conv_backward_input=20, pyramid_deform_conv_backward_parameters=20)
    for name, n in want.items():
        assert len(inspect.signature(getattr(m, name)).parameters) == n, name
