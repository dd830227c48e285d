"""The fused layer nodes of the tri-plane decoder / discriminator (3dgp_b200/torch_utils/ops/modconv.py) against the ORACLE on a machine without a GPU.

Two layers of verification meet at the C ABI: the `-m gpu` tests hold the kernels to the header's contract and to the reference goldens; here the
contract is restated in numpy (tests/abi_emulator.py, host pointers of CPU tensors behind the same ctypes call sites) and the node's host logic --
operand layouts, the padding of the FIR after the stride-2 transposed conv and of its adjoint, tap lists of the input- and weight-gradient launches,
channel padding of the gradient operands, which tensors are saved / rebuilt -- is compared with the reference restatement differentiated by autograd
(oracle/restated.py with DIFFERENTIABLE = True: networks_stylegan2.py:31-88, 128-145, 168-172; layers.py:228-241)."""
import contextlib
import importlib

import numpy as np
import pytest
import torch

import abi_emulator as emu
from oracle import restated as R

tc = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
modconv = importlib.import_module('3dgp_b200.torch_utils.ops.modconv')
sg = importlib.import_module('3dgp_b200.training.networks_stylegan2')
layers = importlib.import_module('3dgp_b200.training.layers')
_lib = importlib.import_module('3dgp_b200._lib')


@pytest.fixture(autouse=True)
def emulated_abi(monkeypatch):
    fake = emu.FakeLib()
    monkeypatch.setattr(_lib, 'lib', lambda: fake)
    monkeypatch.setattr(_lib, 'stream_ptr', lambda: None)
    monkeypatch.setattr(torch.cuda, 'device', lambda _d: contextlib.nullcontext())
    monkeypatch.setattr(tc, 'split_bf16', emu.split)
    monkeypatch.setattr(tc, 'conv_launch', emu.conv_launch_epi)
    monkeypatch.setattr(tc, 'conv_transpose_s2_launch', emu.conv_transpose_s2_launch)
    monkeypatch.setattr(tc, 'wgrad_launch', emu.wgrad_launch)
    monkeypatch.setattr(modconv, 'eligible', lambda x, weight, up, conv_clamp: True)            # the real predicate asks for a CUDA tensor
    monkeypatch.setattr(R, 'DIFFERENTIABLE', True)
    yield
    tc.invalidate_weight_cache()


def _l2rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize('up,noise_mode', [(1, 'random'), (2, 'random'), (2, 'const'), (1, 'none')])
def test_fused_synthesis_layer_forward_and_gradients_vs_oracle(up, noise_mode):
    torch.manual_seed(3)
    C, res, B, wd = 64, 8, 2, 16
    layer = sg.SynthesisLayer(C, C, w_dim=wd, resolution=res, up=up, conv_clamp=None)
    with torch.no_grad():
        layer.noise_strength.fill_(0.3); layer.bias.normal_(0, 0.2)
    x = torch.randn(B, C, res // up, res // up, requires_grad=True)
    w = torch.randn(B, wd, requires_grad=True)
    nz = torch.randn(B, 1, res, res) if noise_mode == 'random' else None
    before = tc.stats['fused']
    y = layer(x, w, noise_mode=noise_mode, fused_modconv=False, noise_in=nz)
    assert tc.stats['fused'] == before + 1, 'the layer did not take the fused node'
    probe = torch.randn_like(y)
    params = [layer.weight, layer.bias, layer.noise_strength, layer.affine.weight, layer.affine.bias]
    got = torch.autograd.grad((y * probe).sum(), [x, w] + params, allow_unused=True)

    sd = {k: v.detach().clone() for k, v in layer.state_dict().items()}
    for k in ('weight', 'bias', 'noise_strength', 'affine.weight', 'affine.bias'):
        sd[k].requires_grad_(True)
    xr, wr = x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True)
    yr = R.synthesis_layer(sd, '', xr, wr, up=up, noise_mode=noise_mode, noise_in=nz, fused_modconv=False)
    ref = torch.autograd.grad((yr * probe).sum(), [xr, wr] + [sd[k] for k in ('weight', 'bias', 'noise_strength', 'affine.weight', 'affine.bias')], allow_unused=True)
    assert _l2rel(y.detach(), yr.detach()) < 1e-5
    for name, a, b in zip(('x', 'w', 'weight', 'bias', 'noise_strength', 'affine.weight', 'affine.bias'), got, ref):
        if b is None or float(b.abs().max()) == 0.0:
            assert a is None or float(a.abs().max()) < 1e-6, name
            continue
        assert _l2rel(a, b) < 2e-4, (name, _l2rel(a, b))       # the emulated backward kernels emit bf16 (hi, lo) pairs: ~2^-16


def test_fused_torgb_forward_and_gradients_vs_oracle():
    torch.manual_seed(4)
    C, B, wd = 64, 2, 16
    layer = sg.ToRGBLayer(C, 96, w_dim=wd)                  # 96 output channels: the gradient operand is zero-padded to 128
    with torch.no_grad():
        layer.bias.normal_(0, 0.2)
    x = torch.randn(B, C, 6, 6, requires_grad=True); w = torch.randn(B, wd, requires_grad=True)
    y = layer(x, w, fused_modconv=False)
    probe = torch.randn_like(y)
    params = [layer.weight, layer.bias, layer.affine.weight, layer.affine.bias]
    got = torch.autograd.grad((y * probe).sum(), [x, w] + params)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in layer.state_dict().items()}
    xr, wr = x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True)
    yr = R.torgb_layer(sd, '', xr, wr, fused_modconv=False)
    ref = torch.autograd.grad((yr * probe).sum(), [xr, wr] + [sd[k] for k in ('weight', 'bias', 'affine.weight', 'affine.bias')])
    assert _l2rel(y.detach(), yr.detach()) < 1e-5
    for name, a, b in zip(('x', 'w', 'weight', 'bias', 'affine.weight', 'affine.bias'), got, ref):
        assert _l2rel(a, b) < 2e-4, (name, _l2rel(a, b))


@pytest.mark.parametrize('k,act,hyper,clamp,terms', [(3, 'lrelu', True, 256, 3), (1, 'linear', False, None, 3), (3, 'lrelu', False, 0.7, 16), (5, 'lrelu', False, None, 3)])
def test_fused_conv2d_layer_forward_and_gradients_vs_oracle(monkeypatch, k, act, hyper, clamp, terms):
    """Conv2dLayer (layers.py:228-241) as the first-order node _ConvBiasAct: hyper-modulation in the operand split, bias / activation / gain / clamp in
    the conv epilogue, clamp-aware activation backward.  `terms` only selects operand formats, which the float64 emulation ignores: the algebra
    (which operands are saved, re-split or re-used by the weight gradient) differs per code and is what is under test."""
    monkeypatch.setattr(modconv, 'conv_act_eligible', lambda *a, **kw: a[2] in (1, 3, 5))
    gradfix = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix')
    torch.manual_seed(5)
    C, B, cd = 64, 2, 12
    layer = layers.Conv2dLayer(C, C, kernel_size=k, activation=act, conv_clamp=clamp, c_dim=cd if hyper else 0, hyper_mod=hyper)
    with torch.no_grad():
        layer.bias.normal_(0, 0.3)
    x = torch.randn(B, C, 6, 5, requires_grad=True)
    c = torch.randn(B, cd, requires_grad=True) if hyper else None
    before = tc.stats['fused']
    with gradfix.tc_terms(terms):
        y = layer(x, c, gain=0.5)
        probe = torch.randn_like(y)
        ins = [x] + ([c] if hyper else []) + list(layer.parameters())
        got = torch.autograd.grad((y * probe).sum(), ins)
    assert tc.stats['fused'] == before + 1
    sd = {kk: v.detach().clone() for kk, v in layer.state_dict().items()}
    names = [n for n, _ in layer.named_parameters()]
    for n in names:
        sd[n].requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    cr = c.detach().clone().requires_grad_(True) if hyper else None
    yr = R.conv2d_layer(sd, '', xr, activation=act, gain=0.5, conv_clamp=clamp, c=cr)
    ref = torch.autograd.grad((yr * probe).sum(), [xr] + ([cr] if hyper else []) + [sd[n] for n in names])
    assert _l2rel(y.detach(), yr.detach()) < 1e-5
    if clamp is not None and clamp < 1:
        assert float((yr.detach().abs() >= clamp * 0.5 - 1e-6).float().mean()) > 0.02, 'the clamp never engaged: the case does not test its mask'
    for name, a, b in zip(['x'] + (['c'] if hyper else []) + names, got, ref):
        assert _l2rel(a, b) < (2e-4 if terms == 3 else 1e-2), (name, _l2rel(a, b))      # 16: ONE bf16 gradient operand (2^-9) by design
