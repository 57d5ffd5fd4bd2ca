"""ResNet stem on the library's kernels: conv 7x7/2 (3 -> 64) with the frozen BatchNorm folded in + ReLU
(lsnet_stem_conv7x7s2_bf16) and the 3x3/2 max-pool (lsnet_maxpool3x3s2_nhwc_bf16) --
mmdet/models/backbones/resnet.py:509-520, 619-623.  Forward only: the reference configs freeze the stem
(frozen_stages=1, resnet.py:569-585); a trainable stem stays on the autograd path of the caller."""
import torch

from .. import lib as L


def pack_stem_weight(w, gamma, beta, mean, var, eps):
    """(64, 3, 7, 7) fp32 + frozen BN -> (bf16 [64, 192] with column ky*24 + kx*3 + ch, fp32 shift [64])."""
    s = gamma / torch.sqrt(var + eps)
    wf = (w * s.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).reshape(64, 7, 21)       # [o][ky][kx*3 + ch]
    wp = torch.zeros((64, 8, 24), device=w.device, dtype=torch.float32)
    wp[:, :7, :21] = wf
    return wp.view(64, 192).to(torch.bfloat16).contiguous(), (beta - mean * s).float().contiguous()


def stem_conv(x, wp, bias):
    """x (B,3,H,W) fp32 or bf16, any strides -> relu(conv7x7/2 + bias) as a logical (B,64,Ho,Wo) channels_last bf16 view."""
    assert x.dim() == 4 and x.shape[1] == 3 and x.dtype in (torch.float32, torch.bfloat16)
    B, _, H, W = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty((B, Ho, Wo, 64), device=x.device, dtype=torch.bfloat16)
    sb, sc, sh, sw = x.stride()
    L.call('lsnet_stem_conv7x7s2_bf16', L.ptr(x), L.c_int(int(x.dtype == torch.bfloat16)), L.c_ll(sb), L.c_ll(sc),
           L.c_ll(sh), L.c_ll(sw), L.c_int(B), L.c_int(H), L.c_int(W), L.ptr(wp), L.ptr(bias), L.ptr(out), L.stream())
    return out.permute(0, 3, 1, 2)


def maxpool3x3s2(x):
    """nn.MaxPool2d(3, stride=2, padding=1) over a channels_last bf16 map."""
    from . import gemm_ops as G
    B, H, W, C, ldp = G.nhwc_geom(x)
    assert ldp == C and x.dtype == torch.bfloat16
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.bfloat16)
    L.call('lsnet_maxpool3x3s2_nhwc_bf16', L.ptr(x), L.c_int(B), L.c_int(H), L.c_int(W), L.c_int(C), L.ptr(y), L.stream())
    return y.permute(0, 3, 1, 2)
