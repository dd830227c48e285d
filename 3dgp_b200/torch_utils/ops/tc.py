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
