"""Per-layer timing of the ResNet-50 trunk convolutions at the bench shape (B=4, 800x1344): the library's own tcgen05
kernels (forward / input gradient / weight gradient) beside cuDNN / cuBLAS through torch on the same bf16 channels_last
tensors.  CUDA events, median of 7, L2 flushed between iterations.  usage: python tools/bench_trunk.py [out.md]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from lsnet_b200.ops.conv import conv2d_packed, conv_out_hw

dev = 'cuda'
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def _queued(fn, n):
    """GPU time of n back-to-back calls: a long spin kernel runs first so that the CPU enqueues everything (Python / ctypes
    launch overhead of a call is ~50-100 us, more than most of these kernels take) before the first call starts."""
    torch.cuda.synchronize()
    torch.cuda._sleep(int(4e7))
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n):
        flush.zero_()
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


T_FLUSH = None


def timeit(fn, n=10):
    global T_FLUSH
    if T_FLUSH is None:
        _queued(lambda: None, n)
        T_FLUSH = min(_queued(lambda: None, n) for _ in range(3))
    fn()
    return min(_queued(fn, n) for _ in range(3)) - T_FLUSH


# (name, I, O, H, W, k, stride, count per step fwd, trainable)
B = 4
SHAPES = [('l1 1x1 64>64', 64, 64, 200, 336, 1, 1, 1, False), ('l1 3x3 64>64', 64, 64, 200, 336, 3, 1, 3, False),
          ('l1 1x1 64>256', 64, 256, 200, 336, 1, 1, 4, False), ('l1 1x1 256>64', 256, 64, 200, 336, 1, 1, 2, False),
          ('l2 1x1 256>128 @200', 256, 128, 200, 336, 1, 1, 1, True), ('l2 3x3s2 128', 128, 128, 200, 336, 3, 2, 1, True),
          ('l2 ds 1x1s2 256>512', 256, 512, 200, 336, 1, 2, 1, True), ('l2 1x1 128>512', 128, 512, 100, 168, 1, 1, 4, True),
          ('l2 1x1 512>128', 512, 128, 100, 168, 1, 1, 3, True), ('l2 3x3 128', 128, 128, 100, 168, 3, 1, 3, True),
          ('l3 1x1 512>256 @100', 512, 256, 100, 168, 1, 1, 1, True), ('l3 3x3s2 256', 256, 256, 100, 168, 3, 2, 1, True),
          ('l3 ds 1x1s2 512>1024', 512, 1024, 100, 168, 1, 2, 1, True), ('l3 1x1 256>1024', 256, 1024, 50, 84, 1, 1, 6, True),
          ('l3 1x1 1024>256', 1024, 256, 50, 84, 1, 1, 5, True), ('l3 3x3 256', 256, 256, 50, 84, 3, 1, 5, True),
          ('l4 1x1 1024>512 @50', 1024, 512, 50, 84, 1, 1, 1, True), ('l4 3x3s2 512', 512, 512, 50, 84, 3, 2, 1, True),
          ('l4 ds 1x1s2 1024>2048', 1024, 2048, 50, 84, 1, 2, 1, True), ('l4 1x1 512>2048', 512, 2048, 25, 42, 1, 1, 3, True),
          ('l4 1x1 2048>512', 2048, 512, 25, 42, 1, 1, 2, True), ('l4 3x3 512', 512, 512, 25, 42, 3, 1, 2, True)]
only = sys.argv[sys.argv.index('--only') + 1] if '--only' in sys.argv else None
if only:
    SHAPES = [s_ for s_ in SHAPES if s_[0] == only]
rows = []
tot = dict(own=0.0, lib=0.0)
for name, I, O, H, W, k, s, cnt, train in SHAPES:
    pad = k // 2
    x = torch.randn(B, I, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(O, I, k, k, device=dev) / (I * k * k) ** 0.5)
    wb = w.permute(0, 2, 3, 1).reshape(O, -1).to(torch.bfloat16).contiguous()
    wt = w.permute(1, 2, 3, 0).reshape(I, -1).to(torch.bfloat16).contiguous()
    wl = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(O, device=dev)
    bl = bias.to(torch.bfloat16)
    Ho, Wo = conv_out_hw(H, W, k, k, s, pad, 1)
    flops = 2.0 * B * Ho * Wo * O * I * k * k
    with torch.no_grad():
        t_own = timeit(lambda: conv2d_packed(x, wb, wt, bias, None, (k, k), s, pad, 1, True))
        t_lib = timeit(lambda: torch.ops.aten.cudnn_convolution_relu(x, wl, bl, [s, s], [pad, pad], [1, 1], 1))
    r = [name, f'{flops / 1e9:.1f}', f'{t_own:.3f}', f'{t_lib:.3f}']
    tot['own'] += cnt * t_own; tot['lib'] += cnt * t_lib
    if train:
        gy = torch.randn(B, O, Ho, Wo, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        xo = x.clone().requires_grad_(True)
        wbo = wb.clone().requires_grad_(True)
        holder = {}
        y = conv2d_packed(xo, wbo, wt, None, None, (k, k), s, pad, 1, False, wgrad_holder=holder)
        # own: input gradient only / weight gradient only (autograd.grad on one input at a time)
        t_dg = timeit(lambda: torch.autograd.grad(y, xo, gy, retain_graph=True))
        t_wg_both = t_dg + timeit(lambda: torch.autograd.grad(y, wbo, gy, retain_graph=True, allow_unused=True))
        t_ldg = timeit(lambda: torch.ops.aten.convolution_backward(gy, x, wl, None, [s, s], [pad, pad], [1, 1], False, [0, 0], 1, [True, False, False]))
        t_lwg = timeit(lambda: torch.ops.aten.convolution_backward(gy, x, wl, None, [s, s], [pad, pad], [1, 1], False, [0, 0], 1, [False, True, False]))
        r += [f'{t_dg:.3f}', f'{t_ldg:.3f}', f'{t_wg_both - t_dg:.3f}', f'{t_lwg:.3f}']
        tot['own'] += cnt * t_wg_both; tot['lib'] += cnt * (t_ldg + t_lwg)
    else:
        r += ['-'] * 4
    r.append(str(cnt))
    rows.append(r)
    print(' | '.join(r), flush=True)
out = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith('--') else 'gpurun_out/trunk_layers.md'
with open(out, 'w') as f:
    f.write('# ResNet-50 trunk convolutions, B=4 800x1344: own tcgen05 kernels vs cuDNN/cuBLAS (ms, CUDA events, L2 flushed)\n\n')
    f.write('| layer | GFLOP | own fwd | lib fwd | own dgrad | lib dgrad | own wgrad | lib wgrad | per step |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|\n')
    for r in rows:
        f.write('| ' + ' | '.join(r) + ' |\n')
    f.write(f"\nweighted per step: own {tot['own']:.2f} ms, library {tot['lib']:.2f} ms\n")
print(open(out).read()[-200:])
