"""Raw (non-autograd) wrappers of the tcgen05 GEMM / implicit-GEMM entry points of liblsnet_sm100.so."""
import weakref

import torch

from .. import lib as L


def _check_2d(t, name):
    assert t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1, f'{name}: bf16 [rows, cols] with unit column stride'
    return t.stride(0)


def gemm(a, bw, bias=None, relu=False, out_dtype=torch.bfloat16, out=None):
    """out[M,N] = a[M,K] @ bw[N,K]^T (+bias)(ReLU).  N % 16 == 0, K % 8 == 0."""
    M, K = a.shape
    N = bw.shape[0]
    lda, ldb = _check_2d(a, 'a'), _check_2d(bw, 'bw')
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    assert out.stride(1) == 1 and out.dtype in (torch.bfloat16, torch.float32)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= N
    L.call('lsnet_gemm_bf16', L.ptr(a), L.c_ll(lda), L.ptr(bw), L.c_ll(ldb), L.ptr(out), L.c_ll(out.stride(0)),
           L.c_int(M), L.c_int(N), L.c_int(K), L.ptr(bias), L.c_int(int(relu)), L.c_int(int(out.dtype == torch.float32)),
           L.stream())
    return out


def gemm_tn(a, b, out=None):
    """out[M,N] (fp32) += a[P,M]^T @ b[P,N]."""
    P, M = a.shape
    N = b.shape[1]
    lda, ldb = _check_2d(a, 'a'), _check_2d(b, 'b')
    if out is None:
        out = torch.zeros((M, N), device=a.device, dtype=torch.float32)
    L.call('lsnet_gemm_tn_bf16', L.ptr(a), L.c_ll(lda), L.ptr(b), L.c_ll(ldb), L.ptr(out), L.c_ll(out.stride(0)),
           L.c_int(P), L.c_int(M), L.c_int(N), L.stream())
    return out


def nhwc_geom(x):
    """x: logical (B,C,H,W) tensor whose memory is pixel-major (channels innermost).  Returns (B,H,W,C,ldp)."""
    B, C, H, W = x.shape
    ldp = x.stride(3)
    assert x.stride(1) == 1 and x.stride(2) == W * ldp and (B == 1 or x.stride(0) == H * W * ldp), \
        f'pixel-major (channels_last) layout required, got strides {x.stride()} for shape {tuple(x.shape)}'
    return B, H, W, C, ldp


def as_nhwc(x, dtype=None):
    """Return x (B,C,H,W) with channels_last memory (and dtype), copying only if needed."""
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    B, C, H, W = x.shape
    ok = x.stride(1) == 1 and x.stride(2) == W * x.stride(3) and (B == 1 or x.stride(0) == H * W * x.stride(3)) \
        and x.stride(3) % 8 == 0 and x.stride(3) >= C
    return x if ok else x.contiguous(memory_format=torch.channels_last)


_PACK_CACHE = {}


def cached_pack(w, kind, fn):
    """Packed bf16 copies of a parameter are reused until the parameter changes in place (optimizer step bumps
    ``_version``): the towers share weights across the 5 pyramid levels, so each weight is packed once per step.
    The entry holds a weak reference to the tensor object: ``id()`` values (and even data pointers) are recycled after
    a tensor dies, so identity must be checked on the live object."""
    key = (id(w), kind)
    hit = _PACK_CACHE.get(key)
    if hit is not None and hit[0]() is w and hit[1] == w._version:
        # Shared towers run their pyramid levels on different streams: a hit from another stream than the one that
        # packed must wait for the packing kernels (inside a captured step this becomes a graph dependency).
        if hit[3] is not None:
            cur = torch.cuda.current_stream(w.device)
            if cur != hit[4]:
                cur.wait_event(hit[3])
                for t in (hit[2] if isinstance(hit[2], (tuple, list)) else (hit[2],)):
                    t.record_stream(cur)
        return hit[2]
    p = fn(w.detach())
    ev = st = None
    if w.is_cuda:
        st = torch.cuda.current_stream(w.device)
        ev = torch.cuda.Event()
        ev.record(st)
    _PACK_CACHE[key] = (weakref.ref(w), w._version, p, ev, st)
    if len(_PACK_CACHE) > 4096:        # dead entries of short-lived tensors
        for k in [k for k, v in _PACK_CACHE.items() if v[0]() is None]:
            del _PACK_CACHE[k]
    return p


def pack_conv_weight(w, flip_transpose=False, n_pad=16):
    """(Cout, Cin, kh, kw) fp32 -> bf16 [Npad, kh*kw*Cpad] tap-major / channel-minor, Cin padded to 64 and rows to
    n_pad.  flip_transpose=True packs the weight of the input-gradient convolution: [Cin_pad16, taps(flipped), Cout_pad64]."""
    if flip_transpose:
        w = w.flip(2, 3).transpose(0, 1)
    co, ci, kh, kw = w.shape
    cpad = (ci + 63) // 64 * 64 if kh * kw > 1 else (ci + 7) // 8 * 8
    npad = (co + n_pad - 1) // n_pad * n_pad
    p = torch.zeros((npad, kh * kw, cpad), device=w.device, dtype=torch.bfloat16)
    p[:co, :, :ci] = w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci)
    return p.view(npad, kh * kw * cpad)


def conv2d_nhwc(x, wp, kh, kw, pad, dil=1, bias=None, relu=False, out_dtype=torch.bfloat16, n_valid=None, ldc=None):
    """Stride-1 'same' conv.  x (B,C,H,W) channels_last bf16; wp from pack_conv_weight.  Returns logical
    (B, n_valid, H, W) view over a pixel-major buffer of ldc channels."""
    B, H, W, C, ldp = nhwc_geom(x)
    N = wp.shape[0]
    ldc = ldc or N
    buf = torch.empty((B, H, W, ldc), device=x.device, dtype=out_dtype)
    if kh == 1 and kw == 1:
        a = torch.as_strided(x, (B * H * W, C), (ldp, 1))
        L.call('lsnet_gemm_bf16', L.ptr(a), L.c_ll(ldp), L.ptr(wp), L.c_ll(wp.stride(0)), L.ptr(buf), L.c_ll(ldc),
               L.c_int(B * H * W), L.c_int(N), L.c_int(C), L.ptr(bias), L.c_int(int(relu)),
               L.c_int(int(out_dtype == torch.float32)), L.stream())
    else:
        L.call('lsnet_conv2d_nhwc_bf16', L.ptr(x), L.c_int(B), L.c_int(H), L.c_int(W), L.c_int(C), L.c_ll(ldp),
               L.ptr(wp), L.c_int(N), L.c_int(kh), L.c_int(kw), L.c_int(pad), L.c_int(pad), L.c_int(dil), L.c_int(dil),
               L.ptr(buf), L.c_ll(ldc), L.ptr(bias), L.c_int(int(relu)), L.c_int(int(out_dtype == torch.float32)),
               L.stream())
    n_valid = n_valid or N
    return buf.permute(0, 3, 1, 2)[:, :n_valid]


def direct_grad(weight):
    """The [Cout, taps*Cin] fp32 view of ``weight``'s gradient memory that a weight-gradient GEMM may accumulate into
    directly (GraphTrainer lays such parameters out tap-major inside its flat gradient buffer), or None."""
    g2 = getattr(weight, '_lsnet_grad2d', None)
    if g2 is not None and weight.grad is not None and weight.grad.data_ptr() == g2.data_ptr():
        return g2
    return None


def conv2d_wgrad_nhwc(dy, x, kh, kw, pad, dil=1, out=None):
    """dW (N, kh*kw, C) fp32 of a stride-1 'same' conv; dy (B,N,H,W), x (B,C,H,W) channels_last bf16, N%8 == C%8 == 0.
    ``out``: accumulate into this [N, kh*kw*C] fp32 buffer instead of a fresh zero-filled one."""
    B, H, W, C, ldx = nhwc_geom(x)
    _, _, _, N, ldy = nhwc_geom(dy)
    assert N % 8 == 0 and C % 8 == 0
    dw = torch.zeros((N, kh * kw, C), device=x.device, dtype=torch.float32) if out is None else out
    if kh == 1 and kw == 1:
        a = torch.as_strided(dy, (B * H * W, N), (ldy, 1))
        b = torch.as_strided(x, (B * H * W, C), (ldx, 1))
        L.call('lsnet_gemm_tn_bf16', L.ptr(a), L.c_ll(ldy), L.ptr(b), L.c_ll(ldx), L.ptr(dw), L.c_ll(C),
               L.c_int(B * H * W), L.c_int(N), L.c_int(C), L.stream())
    else:
        L.call('lsnet_conv2d_wgrad_nhwc_bf16', L.ptr(dy), L.c_ll(ldy), L.ptr(x), L.c_ll(ldx), L.c_int(B), L.c_int(H),
               L.c_int(W), L.c_int(C), L.c_int(N), L.c_int(kh), L.c_int(kw), L.c_int(pad), L.c_int(pad),
               L.c_int(dil), L.c_int(dil), L.ptr(dw), L.stream())
    return dw


def pad_channels_nhwc(t, mult=8, dtype=torch.bfloat16):
    """(B,C,H,W) any layout -> channels_last buffer with C rounded up to `mult` (zero filled); returns the padded
    logical (B,Cpad,H,W) tensor."""
    B, C, H, W = t.shape
    Cp = (C + mult - 1) // mult * mult
    if Cp == C:
        return as_nhwc(t, dtype)
    buf = torch.zeros((B, H, W, Cp), device=t.device, dtype=dtype)
    buf[..., :C] = t.permute(0, 2, 3, 1)
    return buf.permute(0, 3, 1, 2)


def direct_vec(param):
    """The fp32 gradient memory of a 1-D parameter (bias, norm affine) that kernels may ADD into directly -- GraphTrainer
    keeps every gradient as a view of its flat buffer, zeroed once per step -- or None."""
    if param is None or not getattr(param, '_lsnet_direct_vec', False):
        return None
    g = param.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.dim() != 1:
        return None
    return g


def grad_prep(gy, relu_out=None, want_colsum=False, mult=8, colsum_into=None):
    """Stage an upstream gradient (B,C,H,W) for the backward GEMMs in ONE pass: bf16 pixel-major with C padded to
    ``mult``, ReLU-masked by ``relu_out`` (the forward output) if given, plus the per-channel sum (bias gradient).
    Returns (gy_bf16 (B,Cpad,H,W) view, colsum or None).  A gradient that is already bf16 / pixel-major / aligned and
    needs neither mask nor sum is passed through untouched.  ``colsum_into`` (fp32 [C], e.g. the bias parameter's gradient
    memory): the sums are ADDED there by the kernel and None is returned in their place."""
    B, C, H, W = gy.shape
    fn = 'lsnet_grad_prep'
    if colsum_into is not None and want_colsum:
        assert colsum_into.numel() == C and colsum_into.dtype == torch.float32
        fn = 'lsnet_grad_prep_acc'
    else:
        colsum_into = None
    Cp = (C + mult - 1) // mult * mult
    pix_major = gy.stride(1) == 1 and gy.stride(2) == W * gy.stride(3) and (B == 1 or gy.stride(0) == H * W * gy.stride(3))
    if relu_out is None and not want_colsum and pix_major and gy.dtype == torch.bfloat16 and Cp == C \
            and gy.stride(3) % 8 == 0:
        return gy, None
    if relu_out is None and pix_major and gy.dtype == torch.bfloat16 and Cp == C and gy.stride(3) % 8 == 0 \
            and (C // 8) & (C // 8 - 1) == 0 and C // 8 <= 256 and gy.data_ptr() % 16 == 0:
        # already in the GEMM layout: only the bias gradient is missing -> read-only column-sum pass
        colsum = colsum_into if colsum_into is not None else torch.empty(C, device=gy.device, dtype=torch.float32)
        L.call(fn, L.ptr(gy), L.c_int(0), L.c_ll(gy.stride(3)), L.ptr(None), L.c_ll(0), L.c_ll(B * H * W),
               L.c_int(C), L.c_int(Cp), L.ptr(None), L.c_ll(0), L.ptr(colsum), L.stream())
        return gy, (None if colsum_into is not None else colsum)
    if not pix_major or gy.dtype not in (torch.float32, torch.bfloat16):
        gy = gy.contiguous(memory_format=torch.channels_last)
        if gy.dtype not in (torch.float32, torch.bfloat16):
            gy = gy.float()
    ldo = 0
    if relu_out is not None:
        assert relu_out.dtype == torch.bfloat16
        ldo = nhwc_geom(relu_out)[4]
    out = torch.empty((B, H, W, Cp), device=gy.device, dtype=torch.bfloat16)
    colsum = (colsum_into if colsum_into is not None else torch.empty(C, device=gy.device, dtype=torch.float32)) \
        if want_colsum else None
    L.call(fn, L.ptr(gy), L.c_int(int(gy.dtype == torch.float32)), L.c_ll(gy.stride(3)), L.ptr(relu_out),
           L.c_ll(ldo), L.c_ll(B * H * W), L.c_int(C), L.c_int(Cp), L.ptr(out), L.c_ll(Cp), L.ptr(colsum), L.stream())
    return out.permute(0, 3, 1, 2), (None if colsum_into is not None else colsum)
