"""Host logic of the tensor-core convolution wrappers (3dgp_b200/torch_utils/ops/tc.py) on a machine WITHOUT a GPU.

Every conv form of the reference's conv2d_resample (src/torch_utils/ops/conv2d_resample.py:93-141) and every gradient conv2d_gradfix derives from it
(src/torch_utils/ops/conv2d_gradfix.py:113-166) reaches ONE tap-convolution kernel family through tap lists, traversal strides, output lattices and
weight re-layouts that are computed in Python.  Here the ctypes launchers are replaced by a numpy restatement of the C ABI's contract
(tests/abi_emulator.py, written from include/gp3d_b200.h) and the wrappers are compared with torch.nn.functional in float64: a wrong tap offset,
slab index, flip or lattice phase fails here, before any GPU is involved.  (The kernels themselves are held to the same contract by the `-m gpu` tests.)"""
import importlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import abi_emulator as emu

tc = importlib.import_module('3dgp_b200.torch_utils.ops.tc')


@pytest.fixture(autouse=True)
def emulated_abi(monkeypatch):
    monkeypatch.setattr(tc, 'split_bf16', emu.split)
    monkeypatch.setattr(tc, 'conv_launch', emu.conv_launch)
    monkeypatch.setattr(tc, 'conv_transpose_s2_launch', emu.conv_transpose_s2_launch)
    monkeypatch.setattr(tc, 'wgrad_launch', emu.wgrad_launch)


def _rand(*shape, seed=0):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape)).to(torch.float32)


def _close(a, b):
    a, b = a.to(torch.float64), b.to(torch.float64)
    assert a.shape == b.shape
    assert float((a - b).abs().max()) <= 2e-6 * max(float(b.abs().max()), 1.0)       # float32 tensors in, float64 contraction: storage rounding only


@pytest.mark.parametrize('k', [1, 3, 5])
@pytest.mark.parametrize('H,W', [(8, 8), (5, 11)])
def test_stride1_same_conv_and_its_adjoint(k, H, W):
    x, w = _rand(2, 6, H, W), _rand(4, 6, k, k, seed=1)
    _close(tc.conv2d_forward(x, w, 3), F.conv2d(x.double(), w.double(), padding=k // 2))
    # adjoint form (the input gradient of the conv above): conv_transpose2d with the same weight
    g = _rand(2, 4, H, W, seed=2)
    _close(tc.conv2d_forward(g, w, 3, adjoint=True), F.conv_transpose2d(g.double(), w.double(), padding=k // 2))


@pytest.mark.parametrize('H,W,pad', [(8, 8, 1), (9, 13, 1), (8, 8, 0), (10, 6, 2)])
def test_stride2_conv(H, W, pad):
    x, w = _rand(2, 5, H, W), _rand(3, 5, 3, 3, seed=1)
    _close(tc.conv2d_strided_forward(x, w, 2, pad, 3), F.conv2d(x.double(), w.double(), stride=2, padding=pad))


@pytest.mark.parametrize('H,W', [(4, 4), (3, 7)])
@pytest.mark.parametrize('opad', [(0, 0), (1, 1), (1, 0)])
def test_stride2_transposed_conv_polyphase(H, W, opad):
    x, w = _rand(2, 5, H, W), _rand(5, 3, 3, 3, seed=1)         # conv_transpose2d layout [Cin, Cout, k, k]
    _close(tc.conv_transpose2d_s2_forward(x, w, opad, 3), F.conv_transpose2d(x.double(), w.double(), stride=2, padding=0, output_padding=opad))


def _wgrad_reference(x, w_shape, dy, **kw):
    w = torch.zeros(w_shape, dtype=torch.float64, requires_grad=True)
    op = kw.pop('op')
    (op(x.double(), w, **kw) * dy.double()).sum().backward()
    return w.grad


@pytest.mark.parametrize('k,stride,pad,H,W', [(3, 1, 1, 6, 7), (1, 1, 0, 5, 5), (5, 1, 2, 7, 6), (3, 2, 1, 8, 8), (3, 2, 1, 9, 7), (3, 2, 0, 9, 9)])
def test_weight_gradient_of_conv(k, stride, pad, H, W):
    x = _rand(2, 5, H, W)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    dy = _rand(2, 3, Ho, Wo, seed=3)
    got = tc.conv_wgrad(dy, x, k, 'conv', stride, pad, 3)
    _close(got, _wgrad_reference(x, (3, 5, k, k), dy, op=F.conv2d, stride=stride, padding=pad))


@pytest.mark.parametrize('H,W', [(4, 4), (3, 6)])
def test_weight_gradient_of_stride2_transposed_conv(H, W):
    x = _rand(2, 5, H, W)
    dy = _rand(2, 3, 2 * H + 1, 2 * W + 1, seed=3)
    got = tc.conv_wgrad(dy, x, 3, 'transpose', 2, 0, 3)
    _close(got, _wgrad_reference(x, (5, 3, 3, 3), dy, op=F.conv_transpose2d, stride=2, padding=0))


def test_weight_operands_are_cached_per_parameter_version_and_die_with_it():
    p = torch.nn.Parameter(_rand(4, 6, 3, 3))
    n0 = len(tc._weight_cache)
    a = tc.weight_operands(p, 'fwd', lambda w_: w_.permute(0, 2, 3, 1), 3)
    b = tc.weight_operands(p, 'fwd', lambda w_: w_.permute(0, 2, 3, 1), 3)
    assert a[0] is b[0] and len(tc._weight_cache) == n0 + 1                      # hit
    with torch.no_grad():
        p.add_(1.0)                                                             # in-place update bumps the version counter
    c = tc.weight_operands(p, 'fwd', lambda w_: w_.permute(0, 2, 3, 1), 3)
    assert c[0] is not a[0] and float(((c[0] - a[0]) - 1.0).abs().max()) < 1e-6   # refreshed
    tc.invalidate_weight_cache({id(p)})
    assert len(tc._weight_cache) == n0
    tc.weight_operands(p, 'fwd', lambda w_: w_.permute(0, 2, 3, 1), 3)
    del p, a, b, c
    import gc; gc.collect()
    assert len(tc._weight_cache) == n0                                           # the weak reference's callback removed the entry


def test_precision_code_table():
    assert tc.effective_terms(16, grad=True) == 1 and tc.effective_terms(16) == 16
    assert tc.effective_terms(2, grad=True) == 3 and tc.effective_terms(2) == 2
    for terms in (1, 2, 3, 16):
        x_lo, w_lo, x16, w16 = tc.operand_formats(terms)
        assert x16 == w16                                                        # same-format rule of tcgen05.mma.kind::f16 (profiles/r2_operand_format_probe.txt)


# ---------------------------------------------------------------------------------------------------------------------------------
# conv2d_gradfix on the emulated ABI: first- and second-order gradients (the R1 penalty differentiates the discriminator twice,
# reference loss.py:238-253 / conv2d_gradfix.py:141-166) with every primitive routed to the tensor-core wrappers.

gradfix = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix')


def _double_backward(fn, x, w, seeds):
    x = x.clone().requires_grad_(True); w = w.clone().requires_grad_(True)
    y = fn(x, w)
    r1 = _rand(*y.shape, seed=seeds[0]).to(y.dtype)
    gx, gw = torch.autograd.grad((y * r1).sum(), [x, w], create_graph=True)
    r2, r3 = _rand(*gx.shape, seed=seeds[1]).to(y.dtype), _rand(*gw.shape, seed=seeds[2]).to(y.dtype)
    g2x, g2w = torch.autograd.grad((gx * r2).sum() + (gw * r3).sum(), [x, w])
    return [t.detach() for t in (y, gx, gw, g2x, g2w)]


@pytest.mark.parametrize('transpose,stride,pad,H', [(False, 1, 1, 6), (False, 2, 0, 9), (False, 2, 0, 10), (True, 2, 0, 4)])
def test_conv2d_gradfix_first_and_second_order_through_the_tensor_core_wrappers(transpose, stride, pad, H):
    # stride-2 shapes as conv2d_resample issues them: the FIR carries the padding, the conv itself runs with padding 0 (conv2d_resample.py:106-109, 112-126);
    # H = 10 makes the adjoint need output_padding = 1 (the four polyphase launches on a larger output tensor)
    C = 64                                                   # the smallest channel count the tcgen05 kernels accept (ops.tc.channels_eligible)
    x = _rand(1, C, H, H) * 0.5
    w = _rand(C, C, 3, 3, seed=1) * 0.05
    op = gradfix._conv(transpose, w.shape, (stride, stride), (pad, pad), (0, 0), (1, 1), 1, 3)
    before = dict(gradfix.tc_stats)
    got = _double_backward(lambda a, b: op.apply(a, b, None), x, w, (5, 6, 7))
    assert gradfix.tc_stats['aten'] == before['aten'], 'a primitive fell back to ATen: the tap algebra under test was bypassed'
    assert gradfix.tc_stats['tc'] - before['tc'] == 5           # forward, input gradient, weight gradient, and the second-order term of each
    ref_op = (lambda a, b: F.conv_transpose2d(a, b, stride=stride, padding=pad)) if transpose else (lambda a, b: F.conv2d(a, b, stride=stride, padding=pad))
    ref = _double_backward(ref_op, x.double(), w.double(), (5, 6, 7))
    for name, a, b in zip(('y', 'dx', 'dw', 'd2x', 'd2w'), got, ref):
        err = float((a.double() - b).abs().max()) / max(float(b.abs().max()), 1e-12)
        assert err < 1e-5, (name, err)


def test_no_weight_gradients_scope_restores_the_flag_on_error():
    assert not gradfix.weight_gradients_disabled
    with pytest.raises(ValueError):
        with gradfix.no_weight_gradients():
            assert gradfix.weight_gradients_disabled
            raise ValueError('boom')
    assert not gradfix.weight_gradients_disabled
