"""First-contact diagnostics for the tcgen05 GEMM on a real B200: prints error statistics per shape and, on a
mismatch, where in the tile the error lives (helps localise descriptor / swizzle mistakes)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lsnet_b200.ops as ops

torch.manual_seed(0)
dev = 'cuda'
print(torch.cuda.get_device_name(0))


def report(name, out, ref):
    d = (out.double().cpu() - ref.double().cpu()).abs()
    rel = float(d.max() / (ref.abs().max() + 1e-30))
    print(f'{name}: rel_err={rel:.3e} out_absmax={float(out.abs().max()):.3e} ref_absmax={float(ref.abs().max()):.3e}', flush=True)
    if rel > 1e-2:
        R, Cn = d.shape[-2], d.shape[-1]
        rb = d.reshape(-1, Cn)[:128].reshape(16, 8, Cn).amax(1)      # 8-row groups of the first tile
        cb = rb.reshape(16, -1, 8).amax(2) if Cn % 8 == 0 else rb
        print('  err by (8-row group x 8-col group) of first 128 rows, first 8 col groups:')
        print((cb[:, :8] / (ref.abs().max() + 1e-30)).numpy().round(3))
        print('  sample out', out.flatten()[:8].tolist(), 'ref', ref.flatten()[:8].tolist())
    return rel


ok = True
for (M, N, K) in [(128, 256, 64), (128, 256, 128), (128, 32, 64), (256, 128, 256), (1000, 64, 2304), (16800, 256, 2304)]:
    a = torch.randn(M, K).to(torch.bfloat16)
    b = torch.randn(N, K).to(torch.bfloat16)
    ref = a.float() @ b.float().t()
    try:
        out = ops.gemm(a.to(dev), b.to(dev), None, False, torch.float32)
        torch.cuda.synchronize()
        ok &= report(f'gemm_kmajor M{M} N{N} K{K}', out, ref) < 2e-3
    except Exception as e:
        print('gemm_kmajor FAILED', (M, N, K), repr(e), flush=True)
        ok = False
for (P, M, N) in [(64, 128, 256), (128, 128, 256), (1000, 256, 512), (16800, 256, 2304)]:
    a = torch.randn(P, M).to(torch.bfloat16)
    b = torch.randn(P, N).to(torch.bfloat16)
    ref = a.float().t() @ b.float()
    try:
        out = ops.gemm_tn(a.to(dev), b.to(dev))
        torch.cuda.synchronize()
        ok &= report(f'gemm_tn P{P} M{M} N{N}', out, ref) < 2e-3
    except Exception as e:
        print('gemm_tn FAILED', (P, M, N), repr(e), flush=True)
        ok = False
import torch.nn.functional as F
for (B, C, H, W, N) in [(1, 64, 8, 16, 32), (2, 256, 13, 21, 256)]:
    x = torch.randn(B, C, H, W).to(torch.bfloat16)
    w = (torch.randn(N, C, 3, 3) / (C * 9) ** 0.5).to(torch.bfloat16)
    ref = F.conv2d(x.float(), w.float(), None, 1, 1)
    try:
        out = ops.conv2d_same(x.to(dev), w.float().to(dev), None, padding=1, out_fp32=True)
        torch.cuda.synchronize()
        ok &= report(f'conv3x3 B{B} C{C} {H}x{W} N{N}', out.permute(0, 2, 3, 1).reshape(-1, N),
                     ref.permute(0, 2, 3, 1).reshape(-1, N)) < 2e-3
    except Exception as e:
        print('conv FAILED', repr(e), flush=True)
        ok = False
print('DIAG', 'PASS' if ok else 'FAIL')
