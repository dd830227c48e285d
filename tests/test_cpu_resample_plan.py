"""3dgp_b200/torch_utils/ops/conv2d_resample.py: the stage planner (`plan`) and its execution on the emulated ABI against the oracle's restatement of
src/torch_utils/ops/conv2d_resample.py:46-141 (itself pinned by the network goldens), over every route: 1x1 / 3x3 / 5x5 kernels, up / down factors,
2-D and separable filters, symmetric / ragged / negative padding, grouped weights, both weight orientations, flipped filters.  Output extents, values
and the input gradient."""
import importlib
import zlib

import numpy as np
import pytest
import torch

import abi_emulator as emu
from oracle import restated as R
from util import maxrel


@pytest.fixture(autouse=True)
def emulated(monkeypatch):
    emu.install(monkeypatch)


def _filter(taps):
    up = importlib.import_module('3dgp_b200.torch_utils.ops.upfirdn2d')
    return None if taps is None else up.setup_filter(taps)


GRID = [(k, taps, up, down, pad, groups, fw)
        for k in (1, 3, 5)
        for taps in (None, [1, 3, 3, 1], [1, 2, 4, 2, 1])
        for up, down in ((1, 1), (2, 1), (1, 2), (2, 2), (4, 1))
        for pad in (0, 1, [2, 0, 1, 3], [-1, 2, 0, -1])
        for groups in (1, 2)
        for fw in (True, False)
        if not (taps is None and (up > 1 or down > 1) and k > 1 and up > 2)]


def _ids(c):
    k, taps, up, down, pad, groups, fw = c
    return f'k{k}-f{0 if taps is None else len(taps)}-u{up}d{down}-p{pad}-g{groups}-{"corr" if fw else "conv"}'.replace(' ', '')


@pytest.mark.parametrize('case', GRID[::13], ids=[_ids(c) for c in GRID[::13]])
def test_every_route_matches_the_oracle(case):
    cr = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_resample')
    k, taps, up, down, pad, groups, fw = case
    g = torch.Generator().manual_seed(zlib.crc32(_ids(case).encode()))
    Cin, Cout, H, W = 4, 6, 9, 7
    x = torch.randn(2, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin // groups, k, k, generator=g)
    f = _filter(taps)
    fnp = None if f is None else f.numpy()
    want = R.conv2d_resample(x, w, fnp, up=up, down=down, padding=pad, groups=groups, flip_weight=fw, flip_filter=(k == 3))
    if min(want.shape[2:]) < 1:
        pytest.skip('padding crops the whole image')
    xt = x.clone().requires_grad_(True)
    got = cr.conv2d_resample(xt, w, f, up=up, down=down, padding=pad, groups=groups, flip_weight=fw, flip_filter=(k == 3))
    assert tuple(got.shape) == tuple(want.shape)
    assert maxrel(got.detach().numpy(), want.numpy()) < 1e-5
    # input gradient: <A x, v> = <x, A^T v> against the oracle applied to a second input (the operator is linear in x)
    v = torch.randn(want.shape, generator=g)
    gx, = torch.autograd.grad(got, xt, v)
    u = torch.randn(x.shape, generator=g)
    au = R.conv2d_resample(u, w, fnp, up=up, down=down, padding=pad, groups=groups, flip_weight=fw, flip_filter=(k == 3))
    lhs, rhs = float((au.double() * v.double()).sum()), float((u.double() * gx.double()).sum())
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs))


def test_plan_shapes_of_the_layers_on_the_hot_path():
    cr = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_resample')
    # G's up-sampling 3x3 layer (padding 1, [1,3,3,1] filter): ONE stride-2 transposed contraction with no padding of its own, then the FIR with padding 1
    stages = cr.plan(3, 3, 4, 4, 2, 1, (1, 1, 1, 1), True)
    assert stages == (cr.Contract(stride=2, pad=(0, 0), transposed=True, mirrored=True), cr.Fir(True, 1, 1, (1, 1, 1, 1), 4))
    # D's down-sampling 3x3 layer: the FIR carries the whole padding (2, 2), the contraction is a bare stride-2 one
    assert cr.plan(3, 3, 4, 4, 1, 2, (1, 1, 1, 1), True) == (cr.Fir(True, 1, 1, (2, 2, 2, 2), 1), cr.Contract(2, (0, 0), False, False))
    # D's 1x1 skip: decimate first, contract on a quarter of the pixels
    assert cr.plan(1, 1, 4, 4, 1, 2, (0, 0, 0, 0), True) == (cr.Fir(True, 1, 2, (1, 1, 1, 1), 1), cr.Contract(1, (0, 0), False, False))
    # same-resolution layers: one contraction, nothing else
    assert cr.plan(3, 3, 1, 1, 1, 1, (1, 1, 1, 1), True) == (cr.Contract(1, (1, 1), False, False),)
    assert cr.plan(5, 5, 1, 1, 1, 1, (2, 2, 2, 2), False) == (cr.Contract(1, (2, 2), False, True),)
