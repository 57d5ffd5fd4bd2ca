"""Host logic of the fused LSHead glue (no GPU): the (src, mode) table handed to lsnet_pred_reg_fwd/bwd must select the
same channels as LSHead.get_pred_reg (mmdet/models/dense_heads/lsnet_head.py:372-400) for every task branch."""
import math

import pytest
import torch
import torch.nn.functional as F

from lsnet_b200.ops.headglue import pred_reg_table


def _signed_pairs(t):
    r = t.reshape(t.shape[0], -1, 2, *t.shape[2:])
    val, ind = r.max(dim=2)
    return torch.where(ind == 0, -val, val)


def _reference_get_pred_reg(branch, num_vectors, raw1, raw2, kpts=9):
    if raw2 is not None:
        return torch.cat((_signed_pairs(raw1), raw2), 1)
    r = raw1.reshape(raw1.shape[0], -1, 4, *raw1.shape[2:])
    cts, polys = r[:, -1:], r[:, :-1]
    sel = polys[:, ::math.ceil(num_vectors / (kpts - 1))] if branch == 'segm' else polys[:, 1::2]
    offs = torch.cat([sel, cts], 1)
    return _signed_pairs(offs.reshape(offs.shape[0], -1, *offs.shape[3:]))


@pytest.mark.parametrize('branch,num_vectors,n_out', [('bbox', 4, 28), ('segm', 36, 148), ('pose', 17, 72)])
def test_pred_reg_table_matches_get_pred_reg(branch, num_vectors, n_out):
    torch.manual_seed(0)
    o = torch.randn(2, n_out, 3, 4)
    n_sp, src, mode = pred_reg_table(branch, num_vectors, 9, n_out)
    assert len(src) == len(mode) == 18
    sp = F.softplus(o[:, :n_sp])
    ref = _reference_get_pred_reg(branch, num_vectors, sp, o[:, n_sp:] if branch == 'bbox' else None)
    got = torch.stack([torch.where(sp[:, s] >= sp[:, s + 1], -sp[:, s], sp[:, s + 1]) if m == 0 else o[:, s]
                       for s, m in zip(src, mode)], 1)
    assert torch.equal(got, ref)
    # every source channel feeds at most one offset (the backward kernel's inverse table relies on it)
    used = [c for s, m in zip(src, mode) for c in ((s, s + 1) if m == 0 else (s,))]
    assert len(used) == len(set(used))
