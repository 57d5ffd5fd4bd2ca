"""GroupNorm (+ optional residual add, + optional ReLU) over pixel-major bf16 maps on liblsnet_sm100.so
(lsnet_groupnorm_fwd/bwd): fp32 statistics, bf16 in/out, one pass for statistics and one to apply, both directions."""
import torch
from torch.autograd import Function

from .. import lib as L
from . import gemm_ops as G


class _GroupNorm(Function):

    @staticmethod
    def forward(ctx, x, x2, weight, bias, num_groups, eps, relu, pre_sums=None, bias_sink=None):
        ctx.bias_sink = bias_sink        # bias parameter of the layer that produced x: its gradient = colsum(dx), added here
        x = G.as_nhwc(x, torch.bfloat16)
        B, H, W, C, ldx = G.nhwc_geom(x)
        ldx2 = 0
        if x2 is not None:
            x2 = G.as_nhwc(x2, torch.bfloat16)
            ldx2 = G.nhwc_geom(x2)[4]
        w, b = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        y = torch.empty((B, H, W, C), device=x.device, dtype=torch.bfloat16)
        if pre_sums is not None and x2 is None and pre_sums.numel() == B * num_groups * 3 + 1:
            # the producer's GEMM epilogue already accumulated (sum, sum of squares) per (image, group)
            stats = pre_sums
            L.call('lsnet_groupnorm_fwd_pre', L.ptr(x), L.c_ll(ldx), L.c_int(B), L.c_int(H * W), L.c_int(C),
                   L.c_int(num_groups), L.ptr(w), L.ptr(b), L.c_f(eps), L.c_int(int(relu)), L.ptr(stats), L.ptr(y), L.c_ll(C),
                   L.stream())
        else:
            stats = torch.empty(B * num_groups * 3 + 1, device=x.device, dtype=torch.float64)  # sums | ticket | (mean, rstd)
            L.call('lsnet_groupnorm_fwd', L.ptr(x), L.c_ll(ldx), L.ptr(x2), L.c_ll(ldx2), L.c_int(B), L.c_int(H * W),
                   L.c_int(C), L.c_int(num_groups), L.ptr(w), L.ptr(b), L.c_f(eps), L.c_int(int(relu)), L.ptr(stats),
                   L.ptr(y), L.c_ll(C), L.stream())
        ctx.save_for_backward(x, x2, w, b, stats)
        ctx.cfg = (num_groups, eps, relu)
        ctx.affine = (weight, bias)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        x, x2, w, b, stats = ctx.saved_tensors
        num_groups, eps, relu = ctx.cfg
        B, H, W, C, ldx = G.nhwc_geom(x)
        ldx2 = G.nhwc_geom(x2)[4] if x2 is not None else 0
        gy = G.as_nhwc(gy, torch.bfloat16)
        lddy = G.nhwc_geom(gy)[4]
        bstats = torch.empty(B * num_groups * 3 + 1, device=x.device, dtype=torch.float64)
        dx = torch.empty((B, H, W, C), device=x.device, dtype=torch.bfloat16)
        # affine gradients: added straight into the parameters' gradient memory when the trainer exposes it
        tg, tb = G.direct_vec(ctx.affine[0]), G.direct_vec(ctx.affine[1])
        direct = tg is not None and tb is not None and ctx.needs_input_grad[2] and ctx.needs_input_grad[3]
        dgamma = tg if direct else torch.empty(C, device=x.device, dtype=torch.float32)
        dbeta = tb if direct else torch.empty(C, device=x.device, dtype=torch.float32)
        sink = G.direct_vec(ctx.bias_sink) if (ctx.bias_sink is not None and x2 is None) else None
        if ctx.bias_sink is not None and (sink is None or not direct):
            raise RuntimeError('group_norm_nhwc(bias_sink=...): the gradient memory of the parameters is not exposed')
        if direct:
            L.call('lsnet_groupnorm_bwd_acc', L.ptr(x), L.c_ll(ldx), L.ptr(x2), L.c_ll(ldx2), L.ptr(gy), L.c_ll(lddy),
                   L.c_int(B), L.c_int(H * W), L.c_int(C), L.c_int(num_groups), L.ptr(w), L.ptr(b), L.c_f(eps),
                   L.c_int(int(relu)), L.ptr(stats), L.ptr(bstats), L.ptr(dx), L.c_ll(C), L.ptr(dgamma), L.ptr(dbeta),
                   L.ptr(sink), L.stream())
        else:
            L.call('lsnet_groupnorm_bwd', L.ptr(x), L.c_ll(ldx), L.ptr(x2), L.c_ll(ldx2), L.ptr(gy), L.c_ll(lddy),
                   L.c_int(B), L.c_int(H * W), L.c_int(C), L.c_int(num_groups), L.ptr(w), L.ptr(b), L.c_f(eps),
                   L.c_int(int(relu)), L.ptr(stats), L.ptr(bstats), L.ptr(dx), L.c_ll(C), L.ptr(dgamma), L.ptr(dbeta),
                   L.stream())
        dxv = dx.permute(0, 3, 1, 2)
        if direct:
            dgamma = dbeta = None
        return dxv, (dxv if x2 is not None else None), dgamma, dbeta, None, None, None, None, None


def bias_sink_ok(norm_weight, norm_bias, producer_bias):
    """True when group_norm_nhwc(..., bias_sink=producer_bias) can add the producer's bias gradient from inside its backward:
    the trainer exposes the gradient memory of all three parameters (GraphTrainer's flat buffer)."""
    return producer_bias is not None and all(G.direct_vec(p) is not None for p in (norm_weight, norm_bias, producer_bias)) \
        and norm_weight.requires_grad and norm_bias.requires_grad and producer_bias.requires_grad


def group_norm_nhwc(x, num_groups, weight, bias, eps=1e-5, relu=False, residual=None, pre_sums=None, bias_sink=None):
    """relu?(GroupNorm(x (+ residual))) -> (B,C,H,W) channels_last bf16.  ``pre_sums``: fp64 statistics workspace whose sums
    the producer of x accumulated in its GEMM epilogue (ops.dcn: gn_holder) -- the statistics pass is skipped.
    ``bias_sink`` (only when bias_sink_ok): the bias parameter of the layer that produced x; its gradient (the per-channel sum
    of this norm's dx) is added to its gradient memory by the backward apply kernel -- the producer must then NOT compute it."""
    return _GroupNorm.apply(x, residual, weight, bias, int(num_groups), float(eps), bool(relu), pre_sums, bias_sink)
