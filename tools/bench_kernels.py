"""Per-kernel timing at the BASELINE cfg2 level-0 shapes (B=4, 256 ch, 100x168 -> 67 200 pixels): CUDA events on the
launch stream after warm-up, L2 flushed between iterations (write of a 256 MB buffer), against MEASURED_PEAKS.json.
Also the harness ncu attaches to (`--ncu KERNEL` runs that kernel 3 times only)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lsnet_b200.ops as ops
from lsnet_b200.ops import gemm_ops as G

dev = 'cuda'
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json'))) \
    if os.path.exists('MEASURED_PEAKS.json') else dict(hbm_gbs=6650.0, bf16_tflops=1590.0)
ncu_mode = '--ncu' in sys.argv
only = sys.argv[sys.argv.index('--ncu') + 1] if ncu_mode else (
    sys.argv[sys.argv.index('--only') + 1] if '--only' in sys.argv else None)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if ncu_mode:
        return 0.0
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


B, C, H, W = 4, 256, 100, 168
P = B * H * W
g = torch.Generator().manual_seed(0)
x = torch.randn(B, C, H, W, generator=g).to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last)
off = (torch.randn(B, 18, H, W, generator=g) * 1.5).to(dev).contiguous(memory_format=torch.channels_last)
mask = torch.rand(B, 9, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
w = (torch.randn(256, 256, 3, 3, generator=g) / 48).to(dev)
col = torch.randn(P, 2304, generator=g).to(dev, torch.bfloat16)
wp = torch.randn(256, 2304, generator=g).to(dev, torch.bfloat16)
wpT = torch.randn(2304, 256, generator=g).to(dev, torch.bfloat16)
dy = torch.randn(P, 256, generator=g).to(dev, torch.bfloat16)
out = torch.empty(P, 256, device=dev, dtype=torch.bfloat16)
rows = []


def add(name, ms, work, unit):
    if ncu_mode:
        return
    if unit == 'TFLOP/s':
        ach = work / (ms * 1e-3) / 1e12
        frac = ach / peaks['bf16_tflops']
    else:
        ach = work / (ms * 1e-3) / 1e9
        frac = ach / peaks['hbm_gbs']
    rows.append(dict(kernel=name, ms=ms, achieved=ach, unit=unit, frac_of_measured_peak=frac))
    print(f'{name:46s} {ms:8.3f} ms  {ach:9.1f} {unit}  {100 * frac:5.1f}% of measured peak', flush=True)


if only in (None, 'gemm'):
    ms = timeit(lambda: G.gemm(col, wp, None, False, torch.bfloat16, out=out))
    add('gemm_kmajor<256> DCN fwd  M67200 N256 K2304', ms, 2.0 * P * 256 * 2304, 'TFLOP/s')
if only in (None, 'gemm_dcol'):
    ms = timeit(lambda: G.gemm(dy, wpT, None, False, torch.bfloat16))
    add('gemm_kmajor<256> dCol     M67200 N2304 K256', ms, 2.0 * P * 256 * 2304, 'TFLOP/s')
if only in (None, 'conv'):
    wc = G.pack_conv_weight(w)
    ms = timeit(lambda: G.conv2d_nhwc(x, wc, 3, 3, 1))
    add('gemm_kmajor<256> conv3x3  B4 100x168 256->256', ms, 2.0 * P * 256 * 2304, 'TFLOP/s')
if only in (None, 'wgrad'):
    ms = timeit(lambda: G.gemm_tn(dy, col))
    add('gemm_mnmajor DCN dW       P67200 M256 N2304', ms, 2.0 * P * 256 * 2304, 'TFLOP/s')
if only in (None, 'conv_wgrad'):
    dyn = dy.view(B, H, W, 256).permute(0, 3, 1, 2)
    ms = timeit(lambda: G.conv2d_wgrad_nhwc(dyn, x, 3, 3, 1))
    add('gemm_mnmajor conv dW      B4 100x168 256x9x256', ms, 2.0 * P * 256 * 2304, 'TFLOP/s')
if only in (None, 'im2col'):
    ms = timeit(lambda: ops.dcn_im2col(x, off, mask, H, W, 3, 3, (1, 1), (1, 1), (1, 1), (1.0, 1.0), 1))
    add('dcn_im2col DCNv2          B4 100x168 C256', ms, P * (2.0 * C + 4 * 27 + 2.0 * 9 * C), 'GB/s')
if only in (None, 'col2im'):
    gcol = torch.randn(P, 2304, generator=g).to(dev, torch.bfloat16)
    ms = timeit(lambda: ops.dcn_col2im(gcol, x, off, mask, H, W, 3, 3, (1, 1), (1, 1), (1, 1), (1.0, 1.0), 1))
    add('dcn_col2im DCNv2 bf16 reds B4 100x168 C256', ms, P * (2.0 * 9 * C + 2.0 * C + 2.0 * C + 8 * 27), 'GB/s')
    ms = timeit(lambda: ops.dcn_col2im(gcol, x, off, mask, H, W, 3, 3, (1, 1), (1, 1), (1, 1), (1.0, 1.0), 1, dx_fp32=True))
    add('dcn_col2im DCNv2 fp32 reds B4 100x168 C256', ms, P * (2.0 * 9 * C + 2.0 * C + 4.0 * C + 8 * 27), 'GB/s')
    z = torch.zeros_like(off)
    ms = timeit(lambda: ops.dcn_col2im(gcol, x, z, mask, H, W, 3, 3, (1, 1), (1, 1), (1, 1), (1.0, 1.0), 1))
    add('dcn_col2im DCNv2 zero offsets (bench towers)', ms, P * (2.0 * 9 * C + 2.0 * C + 2.0 * C + 8 * 27), 'GB/s')
if only in (None, 'col2im_pyr'):
    # pyramid DCN adjoint on the level-0 grid: sampling level 1 (scale 1/2) and, on the level-1 grid, level 0 (scale 2)
    gcol = torch.randn(P, 2304, generator=g).to(dev, torch.bfloat16)
    x1 = torch.randn(B, C, 50, 84, generator=g).to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ms = timeit(lambda: ops.dcn_col2im(gcol, x1, off, None, H, W, 3, 3, (1, 1), (1, 1), (1, 1), (0.5, 0.5), 1))
    add('dcn_col2im pyramid 100x168 <- 50x84 (s=0.5)', ms, P * (2.0 * 9 * C + 8 * 18) + B * 50 * 84 * 4.0 * C, 'GB/s')
    P1 = B * 50 * 84
    off1 = (torch.randn(B, 18, 50, 84, generator=g) * 1.5).to(dev).contiguous(memory_format=torch.channels_last)
    gcol1 = torch.randn(P1, 2304, generator=g).to(dev, torch.bfloat16)
    ms = timeit(lambda: ops.dcn_col2im(gcol1, x, off1, None, 50, 84, 3, 3, (1, 1), (1, 1), (1, 1), (2.0, 2.0), 1))
    add('dcn_col2im pyramid 50x84 <- 100x168 (s=2)', ms, P1 * (2.0 * 9 * C + 8 * 18) + P * 4.0 * C, 'GB/s')
if only == 'gemm_shapes':
    # every GEMM / implicit-conv shape of one LSHead level set: where does the gemm_kmajor class lose its efficiency?
    lv = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    tot = 0.0
    for (h, w_) in lv:
        Pl = B * h * w_
        xl = torch.randn(B, 256, h, w_, generator=g).to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last)
        x768 = torch.randn(Pl, 768, generator=g).to(dev, torch.bfloat16)
        coll = torch.randn(Pl, 2304, generator=g).to(dev, torch.bfloat16)
        dyl = torch.randn(Pl, 256, generator=g).to(dev, torch.bfloat16)
        dy32 = torch.randn(B, 32, h, w_, generator=g).to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w256 = G.pack_conv_weight(w)
        w27 = G.pack_conv_weight(torch.randn(27, 256, 3, 3, generator=g).to(dev))
        w27t = G.pack_conv_weight(torch.randn(27, 256, 3, 3, generator=g).to(dev), flip_transpose=True)
        w11 = torch.randn(256, 768, generator=g).to(dev, torch.bfloat16)
        w80 = torch.randn(80, 256, generator=g).to(dev, torch.bfloat16)
        cases = [
            ('DCN fwd      N256  K2304 x12', 12, lambda: G.gemm(coll, wp, None, False, torch.bfloat16), 2.0 * Pl * 256 * 2304),
            ('DCN dcol     N2304 K256  x12', 12, lambda: G.gemm(dyl, wpT, None, False, torch.bfloat16), 2.0 * Pl * 256 * 2304),
            ('conv3x3 256->256 fwd+dgrad x6', 6, lambda: G.conv2d_nhwc(xl, w256, 3, 3, 1), 2.0 * Pl * 256 * 2304),
            ('conv_offset 256->27 fwd   x8', 8, lambda: G.conv2d_nhwc(xl, w27, 3, 3, 1, out_dtype=torch.float32), 2.0 * Pl * 32 * 2304),
            ('conv_offset dgrad 32->256 x8', 8, lambda: G.conv2d_nhwc(dy32, w27t, 3, 3, 1), 2.0 * Pl * 256 * 9 * 64),
            ('1x1 768->256 fwd+dgrad    x4', 4, lambda: G.gemm(x768, w11, None, True, torch.bfloat16), 2.0 * Pl * 256 * 768),
            ('1x1 256->80 out           x6', 6, lambda: G.gemm(dyl, w80, None, False, torch.float32), 2.0 * Pl * 80 * 256),
        ]
        for name, cnt, fn, fl in cases:
            ms = timeit(fn, n=6)
            tot += ms * cnt
            print(f'level {h}x{w_:<4d} {name:32s} {ms * 1e3:8.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s   x{cnt} = {ms * cnt:6.3f} ms', flush=True)
    print('sum over levels (with multiplicities):', round(tot, 3), 'ms')
if only in (None, 'dcn_fwd'):
    # whole forward operator: fused (gather -> smem A tile -> tcgen05) against gather -> HBM columns -> GEMM
    from lsnet_b200 import lib as LL
    cfg = (H, W, 3, 3, (1, 1), (1, 1), (1, 1), (1.0, 1.0), 1)
    fl = 2.0 * P * 256 * 2304
    for fused in (1, 0):
        LL.load().lsnet_dcn_fused_enable(fused)
        tag = 'fused' if fused else 'gather+GEMM'
        ms = timeit(lambda: ops.dcn_forward(x, off, mask, wp, None, *cfg, out=out))
        add(f'dcn_forward {tag:12s} DCNv2 100x168 C256 N256', ms, fl, 'TFLOP/s')
        ms = timeit(lambda: ops.dcn_forward(x, off, mask, wp, None, *cfg, out=out, save_col=True))
        add(f'dcn_forward {tag:12s} + column side output', ms, fl, 'TFLOP/s')
    LL.load().lsnet_dcn_fused_enable(1)
    z = torch.zeros_like(off)
    ms = timeit(lambda: ops.dcn_forward(x, z, mask, wp, None, *cfg, out=out))
    add('dcn_forward fused zero offsets (bench towers)', ms, fl, 'TFLOP/s')
    big = off * 8
    ms = timeit(lambda: ops.dcn_forward(x, big, mask, wp, None, *cfg, out=out))
    add('dcn_forward fused offsets x8 (sigma 12 px)', ms, fl, 'TFLOP/s')
    lv = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
    for li, (h, w_) in enumerate(lv):
        Pl = B * h * w_
        xl = torch.randn(B, 256, h, w_, generator=g).to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ol = (torch.randn(B, 18, h, w_, generator=g) * 1.5).to(dev).contiguous(memory_format=torch.channels_last)
        outl = torch.empty(Pl, 256, device=dev, dtype=torch.bfloat16)
        cl = (h, w_, 3, 3, (1, 1), (1, 1), (1, 1), (1.0, 1.0), 1)
        for fused in (1, 0):
            LL.load().lsnet_dcn_fused_enable(fused)
            ms = timeit(lambda: ops.dcn_forward(xl, ol, None, wp, None, *cl, out=outl), n=6)
            add(f'dcn_forward {"fused" if fused else "unfused":8s} DCNv1 level {li} {h}x{w_}', ms, 2.0 * Pl * 256 * 2304, 'TFLOP/s')
        # pyramid: this level's grid sampling the next-finer / next-coarser map
        for (sh_, sw_) in ([lv[li - 1]] if li else []) + ([lv[li + 1]] if li + 1 < len(lv) else []):
            xs = torch.randn(B, 256, sh_, sw_, generator=g).to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last)
            cp = (h, w_, 3, 3, (1, 1), (1, 1), (1, 1), (sh_ / h, sw_ / w_), 1)
            for fused in (1, 0):
                LL.load().lsnet_dcn_fused_enable(fused)
                ms = timeit(lambda: ops.dcn_forward(xs, ol, None, wp, None, *cp, out=outl), n=6)
                add(f'dcn_forward {"fused" if fused else "unfused":8s} pyramid {h}x{w_} <- {sh_}x{sw_}', ms, 2.0 * Pl * 256 * 2304, 'TFLOP/s')
    LL.load().lsnet_dcn_fused_enable(1)
if only in (None, 'dcn_wgrad'):
    # weight gradient: fused re-gather (no columns in HBM) against the split-K GEMM over saved columns (+ the gather that
    # would have to re-create them)
    from lsnet_b200 import lib as LL
    cfg = (H, W, 3, 3, (1, 1), (1, 1), (1, 1), (1.0, 1.0), 1)
    fl = 2.0 * P * 256 * 2304
    dwb = torch.zeros(256, 2304, device=dev)
    LL.load().lsnet_dcn_fused_enable(1)
    ms = timeit(lambda: ops.dcn_backward_weight(dy, x, off, mask, None, *cfg, out=dwb))
    add('dcn_backward_weight fused re-gather 100x168', ms, fl, 'TFLOP/s')
    LL.load().lsnet_set_deterministic(1)
    ms = timeit(lambda: ops.dcn_backward_weight(dy, x, off, mask, None, *cfg, out=dwb))
    add('dcn_backward_weight fused, deterministic 2-stage', ms, fl, 'TFLOP/s')
    LL.load().lsnet_set_deterministic(0)
    ms = timeit(lambda: ops.dcn_backward_weight(dy, x, off, mask, col, *cfg, out=dwb))
    add('dcn_backward_weight saved columns (gemm_mnmajor)', ms, fl, 'TFLOP/s')
    LL.load().lsnet_dcn_fused_enable(0)
    ms = timeit(lambda: ops.dcn_backward_weight(dy, x, off, mask, None, *cfg, out=dwb))
    add('dcn_backward_weight gather -> HBM -> gemm_mnmajor', ms, fl, 'TFLOP/s')
    LL.load().lsnet_dcn_fused_enable(1)
    for (h, w_) in [(50, 84), (25, 42), (13, 21), (7, 11)]:
        Pl = B * h * w_
        xl = torch.randn(B, 256, h, w_, generator=g).to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ol = (torch.randn(B, 18, h, w_, generator=g) * 1.5).to(dev).contiguous(memory_format=torch.channels_last)
        dyl = torch.randn(Pl, 256, generator=g).to(dev, torch.bfloat16)
        cl = (h, w_, 3, 3, (1, 1), (1, 1), (1, 1), (1.0, 1.0), 1)
        coll = ops.dcn_im2col(xl, ol, None, *cl)
        ms = timeit(lambda: ops.dcn_backward_weight(dyl, xl, ol, None, None, *cl, out=dwb), n=6)
        add(f'dcn_backward_weight fused          {h}x{w_}', ms, 2.0 * Pl * 256 * 2304, 'TFLOP/s')
        ms = timeit(lambda: ops.dcn_backward_weight(dyl, xl, ol, None, coll, *cl, out=dwb), n=6)
        add(f'dcn_backward_weight saved columns  {h}x{w_}', ms, 2.0 * Pl * 256 * 2304, 'TFLOP/s')
if only in (None, 'gn'):
    wgt, bias = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    ms = timeit(lambda: ops.group_norm_nhwc(x, 32, wgt, bias, 1e-5, relu=True))
    add('groupnorm fwd (stats+apply) B4 100x168 C256', ms, P * 3.0 * 2 * C, 'GB/s')
if not only:
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(rows, open('gpurun_out/kernels.json', 'w'), indent=1)
