"""Tensor-core (tcgen05 / TMEM) contraction ops of lib3dgp_b200: plain TN GEMM and NHWC implicit-GEMM convolution."""
import torch

from ... import _lib


def gemm_bf16_tn(A, B, out=None, accumulate=False):
    """D[M,N] (+)= A[M,K] @ B[N,K]^T ; A, B bf16 row-major contiguous, D float32.  M % 128 == N % 128 == K % 64 == 0."""
    L = _lib.lib()
    _lib.require_cuda(A, 'A')
    if A.dtype != torch.bfloat16 or B.dtype != torch.bfloat16:
        raise RuntimeError('gemm_bf16_tn: operands must be bfloat16')
    A = A.contiguous(); B = B.contiguous()
    M, K = A.shape
    N, K2 = B.shape
    if K != K2:
        raise RuntimeError('gemm_bf16_tn: inner dimensions differ')
    if out is None:
        out = torch.empty([M, N], dtype=torch.float32, device=A.device)
        accumulate = False
    with torch.cuda.device(A.device):
        rc = L.gp3d_gemm_bf16_tn(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, 1 if accumulate else 0, _lib.stream_ptr())
    _lib.check(rc, 'gemm_bf16_tn')
    return out


def conv2d_nhwc_bf16(x, w, out=None, accumulate=False):
    """Stride-1 'same' convolution (correlation) as an implicit GEMM.  x: [N,H,W,Cin] bf16 (NHWC contiguous),
    w: [Cout,kh,kw,Cin] bf16 with kh == kw in {1, 3}; returns float32 [N,H,W,Cout]."""
    L = _lib.lib()
    _lib.require_cuda(x, 'x')
    if x.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        raise RuntimeError('conv2d_nhwc_bf16: operands must be bfloat16')
    x = x.contiguous(); w = w.contiguous()
    N, H, W, Cin = x.shape
    Cout, kh, kw, Cin2 = w.shape
    if Cin != Cin2 or kh != kw:
        raise RuntimeError('conv2d_nhwc_bf16: bad weight shape')
    if out is None:
        out = torch.empty([N, H, W, Cout], dtype=torch.float32, device=x.device)
        accumulate = False
    with torch.cuda.device(x.device):
        rc = L.gp3d_conv2d_nhwc_bf16(x.data_ptr(), w.data_ptr(), out.data_ptr(), N, H, W, Cin, Cout, kh, 1 if accumulate else 0, _lib.stream_ptr())
    _lib.check(rc, 'conv2d_nhwc_bf16')
    return out


def split_bf16(x_nhwc, styles=None, want_lo=True):
    """x [N, ..., C] channel-minor contiguous (float32 / float16) -> (hi, lo) bfloat16 with x * styles == hi + lo up to 2^-16.
    styles: optional float32 [N, C] per-sample channel scale fused into the split."""
    L = _lib.lib()
    _lib.require_cuda(x_nhwc, 'x')
    assert x_nhwc.is_contiguous()
    N = x_nhwc.shape[0]; C = x_nhwc.shape[-1]
    HW = x_nhwc.numel() // (N * C)
    hi = torch.empty_like(x_nhwc, dtype=torch.bfloat16)
    lo = torch.empty_like(hi) if want_lo else None
    with torch.cuda.device(x_nhwc.device):
        rc = L.gp3d_split_bf16(x_nhwc.data_ptr(), _lib.dtype_code(x_nhwc), _lib.ptr(styles), hi.data_ptr(), _lib.ptr(lo), N, HW, C, _lib.stream_ptr())
    _lib.check(rc, 'split_bf16')
    return hi, lo


def conv_eligible(N, Cin, H, W, Cout, k, stride, padding, dilation, groups):
    """Shapes the tcgen05 implicit-GEMM conv covers (csrc/conv_tc.cu)."""
    if groups != 1 or tuple(stride) != (1, 1) or tuple(dilation) != (1, 1) or k not in (1, 3) or tuple(padding) != (k // 2, k // 2):
        return False
    if Cin % 64 != 0 or not (Cout % 128 == 0 or Cout in (64, 96)):
        return False
    pow2 = lambda v: v >= 1 and (v & (v - 1)) == 0
    if not (pow2(H) and pow2(W)):
        return False
    TW = min(W, 16); TH = min(H, 128 // TW); TN = 128 // (TW * TH)
    return TW * TH * TN == 128 and N % TN == 0


def conv2d_forward(x, w, terms):
    """x [N,Cin,H,W] (any strides, float32/float16), w [Cout,Cin,k,k]  ->  y [N,Cout,H,W] in x.dtype with channels-last strides.
    terms == 3: error-compensated bf16x3 (fp32-grade);  terms == 1: plain bf16 operands, fp32 accumulate."""
    L = _lib.lib()
    N, Cin, H, W = x.shape
    Cout, _, k, _ = w.shape
    xn = x.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)          # NHWC view, contiguous
    wn = w.to(torch.float32).permute(0, 2, 3, 1).contiguous()                            # [Cout,k,k,Cin]
    xh, xl = split_bf16(xn, want_lo=(terms == 3))
    wh, wl = split_bf16(wn, want_lo=(terms == 3))
    y = torch.empty([N, H, W, Cout], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        if terms == 3:
            rc = L.gp3d_conv2d_nhwc_bf16x3(xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr(), y.data_ptr(), N, H, W, Cin, Cout, k, 0, _lib.stream_ptr())
        else:
            rc = L.gp3d_conv2d_nhwc_bf16(xh.data_ptr(), wh.data_ptr(), y.data_ptr(), N, H, W, Cin, Cout, k, 0, _lib.stream_ptr())
    _lib.check(rc, 'conv2d_nhwc_bf16x3' if terms == 3 else 'conv2d_nhwc_bf16')
    y = y.permute(0, 3, 1, 2)
    return y if x.dtype == torch.float32 else y.to(x.dtype)
