"""The product's op wrappers on the product's REAL plugin objects (3dgp_b200/torch_utils/custom_ops.py: argument checks, dtype codes, stride / bias-step
marshalling, output allocation, layout propagation) -- the exact Python path a GPU run takes down to the ctypes call -- with the five C-ABI entry points
served from host memory (tests/abi_emulator.py::install_plugin_library), against the reference goldens.  tests/test_cpu_ops_host.py covers the wrappers on
plugin-level stand-ins; this file adds the marshalling layer in between."""
import importlib
import os

import numpy as np
import pytest
import torch

import abi_emulator as emu
from conftest import ROOT
from oracle import cases
from util import maxrel


@pytest.fixture(autouse=True)
def real_plugins(monkeypatch):
    co = emu.install_plugin_library(monkeypatch)
    for name, plugin in (('upfirdn2d', 'upfirdn2d_plugin'), ('bias_act', 'bias_act_plugin'), ('filtered_lrelu', 'filtered_lrelu_plugin')):
        m = importlib.import_module('3dgp_b200.torch_utils.ops.' + name)
        monkeypatch.setattr(m, '_plugin', co.get_plugin(plugin))
        monkeypatch.setattr(m, '_init', lambda: True)
        assert type(m._plugin).__module__ == '3dgp_b200.torch_utils.custom_ops'


def _gold(name):
    return np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))


@pytest.mark.parametrize('name,kw', cases.bias_act_cases(), ids=[c[0] for c in cases.bias_act_cases()])
@pytest.mark.parametrize('layout', ['dense', 'channels_last'])
def test_bias_act_through_the_plugin_object(name, kw, layout):
    ba = importlib.import_module('3dgp_b200.torch_utils.ops.bias_act')
    x, b = cases.bias_act_inputs(name, kw)
    if layout == 'channels_last' and (x.ndim != 4 or kw['dim'] != 1):
        pytest.skip('channels-last is a 4-D layout')
    g = _gold('bias_act')
    xt = torch.from_numpy(x)
    if layout == 'channels_last':
        xt = xt.contiguous(memory_format=torch.channels_last)      # bias step 1, bias size C: the (i / stepB) % sizeB indexing of the ABI
    xt.requires_grad_(True)
    bt = torch.from_numpy(b).requires_grad_(True) if b is not None else None
    y = ba.bias_act(xt, bt, dim=kw['dim'], act=kw['act'], alpha=kw.get('alpha'), gain=kw.get('gain'), clamp=kw.get('clamp'))
    assert maxrel(y.detach().numpy(), g[name + '/y']) < 1e-5
    if layout == 'channels_last' and x.shape[1] > 1:
        assert y.stride(1) == 1
    gr = torch.autograd.grad(y, [xt] + ([bt] if bt is not None else []), torch.from_numpy(cases.cotangent(y.shape, 11)))
    assert maxrel(gr[0].numpy(), g[name + '/dx']) < 1e-5
    if bt is not None:
        assert maxrel(gr[1].numpy(), g[name + '/db']) < 1e-5


@pytest.mark.parametrize('name,kw', cases.upfirdn2d_cases(), ids=[c[0] for c in cases.upfirdn2d_cases()])
@pytest.mark.parametrize('layout', ['nchw', 'channels_last'])
def test_upfirdn2d_through_the_plugin_object(name, kw, layout):
    up = importlib.import_module('3dgp_b200.torch_utils.ops.upfirdn2d')
    x, f = cases.upfirdn2d_inputs(name, kw)
    g = _gold('upfirdn2d')[name]
    xt = torch.from_numpy(x)
    if layout == 'channels_last':
        xt = xt.contiguous(memory_format=torch.channels_last)
    y = up.upfirdn2d(xt, None if f is None else torch.from_numpy(f), up=kw['up'], down=kw['down'], padding=kw['padding'], flip_filter=kw['flip_filter'], gain=kw['gain'])
    assert tuple(y.shape) == g.shape
    assert np.array_equal(y.numpy(), g) if kw.get('integer', False) else maxrel(y.numpy(), g) < 1e-5
    if layout == 'channels_last' and x.shape[1] > 1:
        assert y.stride(1) == 1


def test_half_precision_tensors_carry_dtype_code_1():
    up = importlib.import_module('3dgp_b200.torch_utils.ops.upfirdn2d')
    ba = importlib.import_module('3dgp_b200.torch_utils.ops.bias_act')
    x = torch.from_numpy(cases.cotangent((1, 3, 8, 8), 3)).half()
    f = up.setup_filter([1, 3, 3, 1])
    y = up.upsample2d(x, f)
    assert y.dtype == torch.float16 and tuple(y.shape) == (1, 3, 16, 16)
    assert maxrel(y.float().numpy(), up.upsample2d(x.float(), f).numpy()) < 2e-3
    z = ba.bias_act(x, torch.zeros(3).half(), act='lrelu')
    assert z.dtype == torch.float16 and maxrel(z.float().numpy(), ba.bias_act(x.float(), torch.zeros(3), act='lrelu').numpy()) < 2e-3
