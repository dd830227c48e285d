"""GPU suite for the tcgen05 / TMEM contraction engine (csrc/gemm_tc.cu, conv_tc.cu)."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


def _tc():
    return importlib.import_module('3dgp_b200.torch_utils.ops.tc')


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (128, 128, 256), (256, 384, 512), (1024, 1024, 2048)])
def test_gemm_bf16_tn_matches_fp32_matmul_of_the_same_bf16_values(M, N, K):
    tc = _tc()
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    A = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    B = torch.randn(N, K, device='cuda', generator=g).bfloat16()
    D = tc.gemm_bf16_tn(A, B)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = A.float() @ B.float().t()
    err = (D - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, err       # products of bf16 values are exact in fp32; only the accumulation order differs
    D2 = tc.gemm_bf16_tn(A, B, out=D.clone(), accumulate=True)
    assert (D2 - 2 * ref).abs().max().item() / ref.abs().max().item() < 2e-5


def test_gemm_rejects_bad_shapes():
    tc = _tc()
    A = torch.zeros(100, 64, device='cuda', dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        tc.gemm_bf16_tn(A, A)


@pytest.mark.parametrize('N,H,W,Cin,Cout,ks', [(1, 16, 16, 64, 128, 3), (2, 32, 32, 128, 256, 3), (8, 4, 4, 64, 128, 3), (1, 64, 64, 128, 96, 1),
                                               (2, 8, 8, 192, 128, 3), (1, 128, 128, 128, 128, 3)])
def test_conv2d_nhwc_bf16_matches_aten_conv_of_the_same_bf16_values(N, H, W, Cin, Cout, ks):
    tc = _tc()
    g = torch.Generator(device='cuda').manual_seed(N * H + Cin + Cout)
    x = torch.randn(N, H, W, Cin, device='cuda', generator=g).bfloat16()
    w = (torch.randn(Cout, ks, ks, Cin, device='cuda', generator=g) / (ks * Cin ** 0.5)).bfloat16()
    y = tc.conv2d_nhwc_bf16(x, w)
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), padding=ks // 2).permute(0, 2, 3, 1)
    err = (y - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err


def test_conv2d_unsupported_shapes_fail_loudly():
    tc = _tc()
    x = torch.zeros(1, 8, 8, 4, device='cuda', dtype=torch.bfloat16)
    w = torch.zeros(128, 3, 3, 4, device='cuda', dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        tc.conv2d_nhwc_bf16(x, w)
