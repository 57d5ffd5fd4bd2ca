"""The row arithmetic the CUDA loss kernels execute (lsnet_b200/csrc/loss_math.cuh), run on the HOST through the
library's lsnet_host_* hooks, against the reference outputs in tests/golden/losses.npz.  No GPU needed: this is what
lets the kernels' math be verified on the CPU-only build box.  Tolerance: 1e-4 relative (north_star)."""
import ctypes
import os

import numpy as np
import pytest

import synth
from oracle import lsnet_oracle as O

FP = ctypes.POINTER(ctypes.c_float)
UP = ctypes.POINTER(ctypes.c_ubyte)
TYPE = {'bbox': 0, 'polygon': 1, 'keypoint': 2}


def _fp(a):
    return a.ctypes.data_as(FP)


@pytest.mark.parametrize('lt', ['bbox', 'polygon', 'keypoint'])
def test_cross_iou_row_value_and_grad(lib, golden_dir, lt):
    g = np.load(os.path.join(golden_dir, 'losses.npz'))
    lib.lsnet_host_cross_iou_row.restype = ctypes.c_float
    for seed in range(4):
        r = synth.loss_rows(lt, 500 + seed)
        tgt, sel = O.directional_targets(r['gt'], r['anchor'], r['weight'])
        pred, tgt = r['pred'].numpy(), tgt.numpy()
        sel = sel.numpy().astype(np.uint8)
        w = r['weight'].numpy().mean(1)
        N, D = pred.shape
        total, grads = 0.0, np.zeros_like(pred)
        for n in range(N):
            if w[n] <= 0:
                continue
            gr = np.zeros(D, np.float32)
            anchor = np.ascontiguousarray(r['anchor'].numpy()[n])
            bb = np.ascontiguousarray(r['bbox_gt'].numpy()[n])
            vs = np.ascontiguousarray(r['vs'].numpy()[n])
            val = lib.lsnet_host_cross_iou_row(TYPE[lt], _fp(np.ascontiguousarray(pred[n])),
                                               _fp(np.ascontiguousarray(tgt[n])),
                                               np.ascontiguousarray(sel[n]).ctypes.data_as(UP), D, _fp(anchor), _fp(bb),
                                               _fp(vs), ctypes.c_float(1e-6), ctypes.c_float(0.2), 9, _fp(gr))
            total += w[n] * val
            grads[n] = w[n] * gr
        loss = 1.5 * total / 7.0
        grads *= 1.5 / 7.0
        ref_loss, ref_grad = float(g[f'{lt}.{seed}.loss']), g[f'{lt}.{seed}.grad']
        assert abs(loss - ref_loss) <= 1e-4 * abs(ref_loss)
        assert np.linalg.norm(grads - ref_grad) <= 1e-4 * np.linalg.norm(ref_grad)
        np.testing.assert_allclose(grads, ref_grad, rtol=2e-3, atol=2e-6)


def test_focal_elem(lib, golden_dir):
    g = np.load(os.path.join(golden_dir, 'losses.npz'))
    lib.lsnet_host_focal_elem.restype = ctypes.c_float
    rng = np.random.RandomState(9)
    logits = (rng.randn(200, 80) * 3).astype(np.float32)
    labels = rng.randint(0, 81, 200)
    w = (rng.rand(200) > 0.1).astype(np.float32)
    total, grad = 0.0, np.zeros_like(logits)
    gr = ctypes.c_float()
    for n in range(200):
        for d in range(80):
            v = lib.lsnet_host_focal_elem(ctypes.c_float(logits[n, d]), int(labels[n]), d, ctypes.c_float(2.0),
                                          ctypes.c_float(0.25), ctypes.byref(gr))
            total += w[n] * v
            grad[n, d] = w[n] * gr.value / 13.0
    assert total / 13.0 == pytest.approx(float(g['focal.loss']), rel=1e-5)
    np.testing.assert_allclose(grad, g['focal.grad'], rtol=1e-4, atol=1e-8)


@pytest.mark.parametrize('lt,NP', [('bbox', 5), ('polygon', 37)])
def test_directional_target_row(lib, golden_dir, lt, NP):
    g = np.load(os.path.join(golden_dir, 'losses.npz'))
    r = synth.loss_rows(lt, 777)
    gt, anchor, w = r['gt'].numpy(), r['anchor'].numpy(), r['weight'].numpy()[:, 0]
    T, S = g[f'dirtgt.{lt}.t'], g[f'dirtgt.{lt}.s']
    for n in range(gt.shape[0]):
        t = np.zeros(4 * NP, np.float32)
        s = np.zeros(4 * NP, np.uint8)
        lib.lsnet_host_directional_target_row(_fp(np.ascontiguousarray(gt[n])), NP, _fp(np.ascontiguousarray(anchor[n])),
                                              int(w[n] > 0), _fp(t), s.ctypes.data_as(UP))
        assert np.array_equal(t, T[n])
        assert np.array_equal(s.astype(bool), S[n])
