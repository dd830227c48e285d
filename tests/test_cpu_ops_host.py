"""Python wrappers of the two plugin ops (3dgp_b200/torch_utils/ops/{bias_act, upfirdn2d}.py: autograd Functions, gradient-of-gradient chain, adjoint
padding, separable-filter passes, memory-format handling) on emulated plugins, against the goldens the reference's own `impl='ref'` paths produced
(tests/golden/bias_act.npz: y / dx / db / d2x for nine activations; upfirdn2d.npz: 14 cases).  The emulated plugins restate the kernels' formula table
and index contract (tests/abi_emulator.py); the kernels themselves are held to the same goldens by the GPU suite."""
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
def emulated(monkeypatch):
    emu.install(monkeypatch)


def _gold(name):
    return np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))


@pytest.mark.parametrize('name,kw', cases.bias_act_cases(), ids=[c[0] for c in cases.bias_act_cases()])
def test_bias_act_wrapper_all_gradient_orders_vs_reference_golden(name, kw):
    ba = importlib.import_module('3dgp_b200.torch_utils.ops.bias_act')
    x, b = cases.bias_act_inputs(name, kw)
    g = _gold('bias_act')
    xt = torch.from_numpy(x).requires_grad_(True)
    bt = torch.from_numpy(b).requires_grad_(True) if b is not None else None
    y = ba.bias_act(xt, bt, dim=kw['dim'], act=kw['act'], alpha=kw.get('alpha'), gain=kw.get('gain'), clamp=kw.get('clamp'))
    assert maxrel(y.detach().numpy(), g[name + '/y']) < 1e-5
    dy = torch.from_numpy(cases.cotangent(y.shape, 11))
    gr = torch.autograd.grad(y, [xt] + ([bt] if bt is not None else []), dy, create_graph=True)
    assert maxrel(gr[0].detach().numpy(), g[name + '/dx']) < 1e-5
    if bt is not None:
        assert maxrel(gr[1].detach().numpy(), g[name + '/db']) < 1e-5
    if (name + '/d2x') in g.files and gr[0].requires_grad:
        v = torch.from_numpy(cases.cotangent(y.shape, 12))
        g2 = torch.autograd.grad(gr[0], xt, v, allow_unused=True)[0]
        g2 = torch.zeros_like(xt) if g2 is None else g2
        ref2 = g[name + '/d2x']
        assert np.abs(g2.numpy() - ref2).max() < 1e-5 * max(1.0, np.abs(ref2).max())


@pytest.mark.parametrize('name,kw', cases.upfirdn2d_cases(), ids=[c[0] for c in cases.upfirdn2d_cases()])
def test_upfirdn2d_wrapper_forward_and_adjoint_vs_reference_golden(name, kw):
    up = importlib.import_module('3dgp_b200.torch_utils.ops.upfirdn2d')
    x, f = cases.upfirdn2d_inputs(name, kw)
    g = _gold('upfirdn2d')[name]
    ft = None if f is None else torch.from_numpy(f)
    xt = torch.from_numpy(x).requires_grad_(True)
    y = up.upfirdn2d(xt, ft, up=kw['up'], down=kw['down'], padding=kw['padding'], flip_filter=kw['flip_filter'], gain=kw['gain'])
    assert tuple(y.shape) == g.shape
    if kw.get('integer', False):
        assert np.array_equal(y.detach().numpy(), g)
    else:
        assert maxrel(y.detach().numpy(), g) < 1e-5
    # the backward is upfirdn2d with up / down swapped and the adjoint padding (upfirdn2d.py:250-269): <A x, v> == <x, A^T v>
    v = torch.from_numpy(cases.cotangent(y.shape, 5))
    gx, = torch.autograd.grad(y, xt, v)
    u = torch.from_numpy(cases.cotangent(x.shape, 6))
    yu = up.upfirdn2d(u, ft, up=kw['up'], down=kw['down'], padding=kw['padding'], flip_filter=kw['flip_filter'], gain=kw['gain'])
    lhs, rhs = float((yu.double() * v.double()).sum()), float((u.double() * gx.double()).sum())
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs))


@pytest.mark.parametrize('name,kw', cases.filtered_lrelu_cases(), ids=[c[0] for c in cases.filtered_lrelu_cases()])
def test_filtered_lrelu_wrapper_generic_route_vs_reference_golden(monkeypatch, name, kw):
    """filtered_lrelu.py's autograd Function on a plugin WITHOUT a specialised kernel (return code -1): the generic route and its backward -- the same op
    with up / down swapped, the adjoint padding and the stored sign codes read back at the shifted offsets (filtered_lrelu.py:252-263) -- against y / dx / db
    of the reference's `_filtered_lrelu_ref`."""
    fl = importlib.import_module('3dgp_b200.torch_utils.ops.filtered_lrelu')
    monkeypatch.setattr(fl, '_plugin', emu.FilteredLreluPlugin); monkeypatch.setattr(fl, '_init', lambda: True)
    x, fu, fd, b = cases.filtered_lrelu_inputs(name, kw)
    g = _gold('filtered_lrelu')
    xt = torch.from_numpy(x).requires_grad_(True); bt = torch.from_numpy(b).requires_grad_(True)
    F_ = fl._filtered_lrelu_cuda(up=kw['up'], down=kw['down'], padding=kw['padding'], gain=kw['gain'], slope=kw['slope'], clamp=kw['clamp'], flip_filter=False)
    y = F_.apply(xt, torch.from_numpy(fu), torch.from_numpy(fd), bt, None, 0, 0)        # the public wrapper refuses CPU tensors; this is what it dispatches to
    assert maxrel(y.detach().numpy(), g[name + '/y']) < 1e-5
    dy = torch.from_numpy(cases.cotangent(y.shape, 13))
    gx, gb = torch.autograd.grad(y, [xt, bt], dy)
    assert maxrel(gx.numpy(), g[name + '/dx']) < 2e-5 and maxrel(gb.numpy(), g[name + '/db']) < 2e-5


@pytest.mark.parametrize('taps', [[1, 3, 3, 1], [1, 2, 1], [1, 4, 6, 4, 1], [1, 1, 2, 3, 3, 2, 1, 1]], ids=lambda t: f'f{len(t)}')
@pytest.mark.parametrize('pad', [0, [1, 2], [2, 0, -1, 1]], ids=lambda p: f'p{p}'.replace(' ', ''))
def test_resampling_helpers_and_the_spec_algebra_vs_oracle(taps, pad):
    """filter2d / upsample2d / downsample2d (centring margins of upfirdn2d.py:277-387) against the oracle; FirSpec.out_extent is the extent the plugin
    returns; the adjoint of the adjoint restores the spec whenever the forward pass dropped nothing (exact multiples)."""
    from oracle import restated as R
    up = importlib.import_module('3dgp_b200.torch_utils.ops.upfirdn2d')
    f = up.setup_filter(taps)                              # 8 taps -> separable (two passes), shorter -> 2-D
    fnp = R.setup_filter(taps)
    assert np.array_equal(f.numpy(), fnp)
    x = torch.from_numpy(cases.cotangent((2, 3, 10, 12), 21))
    for name, kw in (('filter2d', {}), ('upsample2d', dict(up=2)), ('upsample2d', dict(up=[2, 1])), ('downsample2d', dict(down=2)), ('downsample2d', dict(down=[1, 2]))):
        got = getattr(up, name)(x, f, padding=pad, flip_filter=True, **kw)
        want = getattr(R, name)(x.numpy(), fnp, padding=pad, flip_filter=True, **kw)
        assert tuple(got.shape) == want.shape, (name, kw)
        assert maxrel(got.numpy(), want) < 1e-5, (name, kw)
    fw, fh = up._get_filter_size(f)
    for u, d in ((1, 1), (2, 1), (1, 2), (3, 2)):
        spec = up._upfirdn2d_cuda(up=u, down=d, padding=pad, flip_filter=False, gain=2)
        y = spec.apply(x, f)
        assert tuple(y.shape[2:]) == spec.out_extent(10, 12, fh, fw)
        adj = spec.adjoint((10, 12), tuple(y.shape[2:]), (fh, fw))
        assert adj.up == spec.down and adj.down == spec.up and adj.flip != spec.flip and adj.out_extent(*y.shape[2:], fh, fw) == (10, 12)
        if d == 1:
            assert adj.adjoint(tuple(y.shape[2:]), (10, 12), (fh, fw)) == spec
