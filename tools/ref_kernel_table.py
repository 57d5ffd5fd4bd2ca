"""The reference's own CUDA kernels (oracle/_ref: mmdet/ops/dcn compiled unmodified for sm_100a, fp32) timed beside
liblsnet_sm100 on the shapes of the hot path: DCNv2 forward and backward on the five head level grids of BASELINE
configs[1] (B=4, 256 ch) and on the grouped X-101 backbone sites (groups=64; C = 512 / 1024 / 2048).  CUDA events, median of
7 after 2 warm-ups, L2 flushed between iterations.  Writes a markdown table (BASELINE.md 2b "kernel to beat").
usage (GPU box): python tools/ref_kernel_table.py [out.md]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import lsnet_b200.ops as ops
from oracle import build_ref

out_path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/r02_reference_kernels.md'
ext = build_ref.load_ext()
dev = 'cuda'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=7):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


rows = []


def case(name, B, C, H, W, groups, stride):
    g = torch.Generator().manual_seed(1)
    Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    x = torch.randn(B, C, H, W, generator=g).to(dev)
    off = (torch.randn(B, 18, Ho, Wo, generator=g) * 1.5).to(dev)
    mask = torch.rand(B, 9, Ho, Wo, generator=g).to(dev)
    w = (torch.randn(C, C // groups, 3, 3, generator=g) / (9 * C // groups) ** 0.5).to(dev)
    b = torch.zeros(C, device=dev)
    gy = torch.randn(B, C, Ho, Wo, generator=g).to(dev)
    e = x.new_empty(0)
    ro = torch.empty(B, C, Ho, Wo, device=dev)
    gi, go, gm, gw, gb = (torch.zeros_like(t) for t in (x, off, mask, w, b))
    t_rf = timeit(lambda: ext.modulated_deform_conv_forward(x, w, b, e, off, mask, ro, e, 3, 3, stride, stride, 1, 1, 1, 1,
                                                            groups, 1, False))
    t_rb = timeit(lambda: ext.modulated_deform_conv_backward(x, w, b, e, off, mask, e, gi, gw, gb, go, gm, gy, 3, 3, stride,
                                                             stride, 1, 1, 1, 1, groups, 1, False))
    xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    offb = off.contiguous(memory_format=torch.channels_last).requires_grad_(True)
    maskb = mask.contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wb = w.clone().requires_grad_(True)
    gyb = gy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    t_of = timeit(lambda: ops.modulated_deform_conv(xb.detach(), offb.detach(), maskb.detach(), wb.detach(), None, stride, 1, 1, groups))

    def fb():
        y = ops.modulated_deform_conv(xb, offb, maskb, wb, None, stride, 1, 1, groups)
        torch.autograd.grad(y, [xb, offb, maskb, wb], gyb)
    t_ofb = timeit(fb)
    t_ob = max(t_ofb - t_of, 1e-6)
    rows.append((name, f'{B}x{C}x{H}x{W}', groups, stride, t_rf, t_rb, t_of, t_ob))
    print(f'{name:28s} ref fwd {t_rf:8.3f} bwd {t_rb:8.3f} ms | ours fwd {t_of:7.3f} bwd {t_ob:7.3f} ms | x{t_rf / t_of:5.1f} / x{t_rb / t_ob:5.1f}', flush=True)


for i, (h, w_) in enumerate([(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]):
    case(f'head level {i}', 4, 256, h, w_, 1, 1)
case('X-101 c3 first block', 4, 512, 200, 336, 64, 2)
case('X-101 c3', 4, 512, 100, 168, 64, 1)
case('X-101 c4 first block', 4, 1024, 100, 168, 64, 2)
case('X-101 c4', 4, 1024, 50, 84, 64, 1)
case('X-101 c5 first block', 4, 2048, 50, 84, 64, 2)
case('X-101 c5', 4, 2048, 25, 42, 64, 1)
os.makedirs(os.path.dirname(out_path) or '.', exist_ok=True)
with open(out_path, 'w') as f:
    f.write('# Reference DCNv2 kernels (mmdet/ops/dcn, sm_100a build, fp32) vs liblsnet_sm100 (bf16) on the hot-path shapes\n\n')
    f.write('CUDA events, median of 7, L2 flushed between iterations, one B200.  Reference = `modulated_deform_conv_forward` / '
            '`_backward` of the unmodified extension in oracle/_ref (per-sample im2col + cuBLAS addmm_, atomics col2im).  Ours = '
            '`ops.modulated_deform_conv` forward and (forward+backward) - forward through the whole-operator C ABI (backward = '
            'dX, dOffset, dMask, dW).\n\n')
    f.write('| site | x (B,C,H,W) | groups | stride | ref fwd ms | ref bwd ms | ours fwd ms | ours bwd ms | fwd speed-up | bwd speed-up |\n')
    f.write('|---|---|---:|---:|---:|---:|---:|---:|---:|---:|\n')
    for n, shp, g_, s_, a, b_, c, d in rows:
        f.write(f'| {n} | {shp} | {g_} | {s_} | {a:.3f} | {b_:.3f} | {c:.3f} | {d:.3f} | {a / c:.1f}x | {b_ / d:.1f}x |\n')
print(open(out_path).read())
