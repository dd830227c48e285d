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


# ---------------------------------------------------------------------------------------------------------------------------------
# One discriminator block (networks_discriminator.py:67-90) incl. its down-sampling layers (conv2d_resample: FIR then stride-2 conv, FIR-decimate then
# 1x1 conv) and the R1-style double backward (loss.py:238-253), on emulated plugins + the emulated tensor-core ABI.

def _oracle_block(sd, x, c, down):
    s = float(np.sqrt(0.5))
    y = R.conv2d_layer(sd, 'skip.', x, down=down, gain=s)
    h = R.conv2d_layer(sd, 'conv0.', x, activation='lrelu')
    h = R.conv2d_layer(sd, 'conv1.', h, activation='lrelu', down=down, gain=s, c=c)
    return y + h


@pytest.mark.parametrize('down,second_order', [(2, False), (1, False), (2, True)])
def test_discriminator_block_first_order_and_r1_double_backward_vs_oracle(monkeypatch, down, second_order):
    nd = importlib.import_module('3dgp_b200.training.networks_discriminator')
    upf = importlib.import_module('3dgp_b200.torch_utils.ops.upfirdn2d')
    bact = importlib.import_module('3dgp_b200.torch_utils.ops.bias_act')
    gradfix = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix')
    monkeypatch.setattr(modconv, 'conv_act_eligible', lambda x, w, k, up, dn, pad, act, cin, cout: up == 1 and dn == 1 and k in (1, 3) and cin % 64 == 0 and cout % 64 == 0)
    monkeypatch.setattr(upf, '_plugin', emu.Upfirdn2dPlugin); monkeypatch.setattr(upf, '_init', lambda: True)
    monkeypatch.setattr(bact, '_plugin', emu.BiasActPlugin); monkeypatch.setattr(bact, '_init', lambda: True)
    # the public wrappers refuse CPU tensors (the product has no CPU path): enter below the guard, at the autograd Functions they dispatch to
    monkeypatch.setattr(upf, 'upfirdn2d', lambda x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda':
                        upf._upfirdn2d_cuda(up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain).apply(x, f))
    monkeypatch.setattr(bact, 'bias_act', lambda x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda':
                        bact._bias_act_cuda(dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp).apply(x, b))
    monkeypatch.setattr(gradfix, 'conv2d', lambda input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1:
                        gradfix._conv(False, weight.shape, gradfix._tuple2(stride), gradfix._tuple2(padding), (0, 0), gradfix._tuple2(dilation), groups,
                                      gradfix._terms_for(input.dtype)).apply(input, weight, bias))
    torch.manual_seed(7)
    C, B, cd = 64, 2, 10
    blk = nd.DiscriminatorBlock(None, C, C, C, resolution=8, img_channels=4, first_layer_idx=2, down=down, c_dim=cd, hyper_mod=True, conv_clamp=None, use_fp16=False)
    with torch.no_grad():
        for n_, p_ in blk.named_parameters():
            if n_.endswith('bias'):
                p_.normal_(0, 0.2)
    x = torch.randn(B, C, 8, 8, requires_grad=True); c = torch.randn(B, cd)
    names = [n for n, _ in blk.named_parameters()]
    sd = {k: v.detach().clone() for k, v in blk.state_dict().items()}
    for n in names:
        sd[n].requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    before = dict(tc.stats)
    if not second_order:
        y = blk(x, None, c=c)
        probe = torch.randn_like(y)
        got = torch.autograd.grad((y * probe).sum(), [x] + list(blk.parameters()), allow_unused=True)      # fromrgb exists but is idle in an inner block
        assert tc.stats['fused'] - before['fused'] == (1 if down == 2 else 3), 'stride-1 layers of the block run as fused first-order nodes'
        yr = _oracle_block(sd, xr, c, down)
        ref = torch.autograd.grad((yr * probe).sum(), [xr] + [sd[n] for n in names], allow_unused=True)
    else:       # R1: gradient of the squared input-gradient norm w.r.t. the weights, weight gradients off inside the inner pass (loss.py:245-249)
        with layers.first_order_only(False):
            y = blk(x, None, c=c)
        assert tc.stats['fused'] == before['fused'], 'the twice-differentiable composition must be used under first_order_only(False)'
        with gradfix.no_weight_gradients():
            gx, = torch.autograd.grad(y.sum(), [x], create_graph=True)
        got = torch.autograd.grad(gx.square().sum(), list(blk.parameters()), allow_unused=True)
        got = [x.grad] + list(got)
        yr = _oracle_block(sd, xr, c, down)
        gxr, = torch.autograd.grad(yr.sum(), [xr], create_graph=True)
        ref = [None] + list(torch.autograd.grad(gxr.square().sum(), [sd[n] for n in names], allow_unused=True))
        assert _l2rel(gx.detach(), gxr.detach()) < 1e-4
    assert tc.stats['aten'] == before['aten'], 'a 64-channel convolution of the block fell back to ATen'
    assert _l2rel(y.detach(), yr.detach()) < 1e-5
    for name, a, b in zip(['x'] + names, got, ref):
        if b is None or float(b.abs().max()) == 0.0:
            assert a is None or float(a.abs().max()) < 1e-6, name
            continue
        assert _l2rel(a, b) < 3e-4, (name, _l2rel(a, b))
