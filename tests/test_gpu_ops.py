"""GPU parity suite for the plugin ops: sm_100a kernels (through the reference-shaped plugin objects and the C ABI)
vs the golden vectors produced by the unmodified reference, and vs the CPU oracle on fresh seeds."""
import importlib

import numpy as np
import pytest
import torch

from oracle import cases, restated as R
from util import maxrel

pytestmark = pytest.mark.gpu
TOL = 1e-5   # fp32 ops: only summation order / fma contraction / fast-math intrinsics differ


def _ops():
    return (importlib.import_module('3dgp_b200.torch_utils.ops.bias_act'),
            importlib.import_module('3dgp_b200.torch_utils.ops.upfirdn2d'),
            importlib.import_module('3dgp_b200.torch_utils.ops.filtered_lrelu'))


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize('name,kw', cases.upfirdn2d_cases(), ids=[c[0] for c in cases.upfirdn2d_cases()])
def test_upfirdn2d_vs_golden(golden, name, kw):
    _, up, _ = _ops()
    x, f = cases.upfirdn2d_inputs(name, kw)
    g = golden('upfirdn2d')[name]
    ft = None if f is None else cu(f)
    y = up.upfirdn2d(cu(x), ft, up=kw['up'], down=kw['down'], padding=kw['padding'], flip_filter=kw['flip_filter'], gain=kw['gain'])
    assert tuple(y.shape) == g.shape                       # integer index formula
    if kw.get('integer', False):
        assert np.array_equal(y.cpu().numpy(), g)          # bit-exact
    else:
        assert maxrel(y.cpu().numpy(), g) < TOL
    # channels-last storage goes through the C-minor kernel and must give the same numbers
    if x.shape[1] % 4 == 0 and f is not None and f.ndim == 2:
        ycl = up.upfirdn2d(cu(x).contiguous(memory_format=torch.channels_last), ft, up=kw['up'], down=kw['down'], padding=kw['padding'],
                           flip_filter=kw['flip_filter'], gain=kw['gain'])
        assert ycl.stride(1) == 1
        assert torch.equal(ycl.contiguous(), y)


def test_upfirdn2d_gradient_is_adjoint():
    _, up, _ = _ops()
    rs = np.random.RandomState(3)
    x = cu(rs.standard_normal((2, 3, 9, 11)).astype(np.float32)).requires_grad_(True)
    f = cu(cases._f2d(cases.F1331))
    y = up.upfirdn2d(x, f, up=2, down=1, padding=[2, 1, 2, 1], gain=4)
    dy = cu(rs.standard_normal(tuple(y.shape)).astype(np.float32))
    (dx,) = torch.autograd.grad(y, x, dy)
    # <A x, dy> == <x, A^T dy>
    lhs = (y * dy).sum().item(); rhs = (x * dx).sum().item()
    assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(lhs))
    xo = x.detach().cpu().numpy()
    yo = R.upfirdn2d(xo, f.cpu().numpy(), up=2, padding=[2, 1, 2, 1], gain=4)
    assert maxrel(y.detach().cpu().numpy(), yo) < TOL


def test_upfirdn2d_fp16_and_large_shape_properties():
    _, up, _ = _ops()
    f = cu(cases._f2d(cases.F1331))
    # BASELINE-size layer: [B,128,256,256] -> up 2 (skip path).  Linearity + DC gain (filter sums to 1, gain 4 = up^2).
    x = torch.ones([2, 128, 256, 256], device='cuda')
    y = up.upsample2d(x, f)
    assert tuple(y.shape) == (2, 128, 512, 512)
    inner = y[:, :, 4:-4, 4:-4]
    assert torch.allclose(inner, torch.ones_like(inner), atol=1e-6)
    a = torch.randn([1, 64, 128, 128], device='cuda'); b = torch.randn_like(a)
    assert torch.allclose(up.upsample2d(a + 2 * b, f), up.upsample2d(a, f) + 2 * up.upsample2d(b, f), atol=1e-4)
    xh = torch.randn([2, 8, 32, 32], device='cuda', dtype=torch.float16)
    yh = up.downsample2d(xh, f)
    yo = R.downsample2d(xh.float().cpu().numpy(), f.cpu().numpy())
    assert yh.dtype == torch.float16 and maxrel(yh.float().cpu().numpy(), yo) < 2e-3


def test_upfirdn2d_errors():
    _, up, _ = _ops()
    with pytest.raises(RuntimeError):
        up.upfirdn2d(torch.zeros([1, 1, 2, 2], device='cuda'), cu(cases._f2d([1, 4, 6, 4, 1])))   # output < 1x1
    with pytest.raises(RuntimeError):
        up._plugin.upfirdn2d(torch.zeros([1, 1, 4, 4], device='cuda'), torch.ones([2, 2], device='cuda', dtype=torch.float16), 1, 1, 1, 1, 0, 0, 0, 0, False, 1.0)


@pytest.mark.parametrize('name,kw', cases.bias_act_cases(), ids=[c[0] for c in cases.bias_act_cases()])
def test_bias_act_vs_golden(golden, name, kw):
    ba, _, _ = _ops()
    x, b = cases.bias_act_inputs(name, kw)
    g = golden('bias_act')
    xt = cu(x).requires_grad_(True)
    bt = cu(b).requires_grad_(True) if b is not None else None
    y = ba.bias_act(xt, bt, dim=kw['dim'], act=kw['act'], alpha=kw.get('alpha'), gain=kw.get('gain'), clamp=kw.get('clamp'))
    assert maxrel(y.detach().cpu().numpy(), g[name + '/y']) < TOL
    dy = cu(cases.cotangent(y.shape, 11))
    gr = torch.autograd.grad(y, [xt] + ([bt] if bt is not None else []), dy, create_graph=True)
    assert maxrel(gr[0].detach().cpu().numpy(), g[name + '/dx']) < 5e-5
    if bt is not None:
        assert maxrel(gr[1].detach().cpu().numpy(), g[name + '/db']) < 5e-5
    if (name + '/d2x') in g.files and gr[0].requires_grad:
        v = cu(cases.cotangent(y.shape, 12))
        g2 = torch.autograd.grad(gr[0], xt, v, allow_unused=True)[0]
        g2 = torch.zeros_like(xt) if g2 is None else g2
        ref2 = g[name + '/d2x']
        assert np.abs(g2.cpu().numpy() - ref2).max() < 5e-5 * max(1.0, np.abs(ref2).max())


def test_bias_act_half_and_channels_last():
    ba, _, _ = _ops()
    x = torch.randn([4, 16, 8, 8], device='cuda')
    b = torch.randn([16], device='cuda')
    y = ba.bias_act(x, b, act='lrelu')
    ycl = ba.bias_act(x.contiguous(memory_format=torch.channels_last), b, act='lrelu')
    assert torch.equal(ycl.contiguous(), y)
    yh = ba.bias_act(x.half(), b.half(), act='lrelu', clamp=256)
    assert maxrel(yh.float().cpu().numpy(), y.cpu().numpy()) < 3e-3
    yb = ba.bias_act(x.bfloat16(), b.bfloat16(), act='lrelu')
    assert maxrel(yb.float().cpu().numpy(), y.cpu().numpy()) < 2e-2


def test_bias_act_full_size_roundtrip():
    """BASELINE-size tensor (one 512^2 x 128 layer of one image): grad of lrelu through saved y is +-gain mask."""
    ba, _, _ = _ops()
    x = torch.randn([1, 128, 512, 512], device='cuda', requires_grad=True)
    y = ba.bias_act(x, None, act='lrelu')
    (dx,) = torch.autograd.grad(y.sum(), x)
    g = float(np.sqrt(2))
    expect = torch.where(x > 0, torch.full_like(x, g), torch.full_like(x, 0.2 * g))
    assert torch.allclose(dx, expect, rtol=1e-6, atol=0)


@pytest.mark.parametrize('name,kw', cases.filtered_lrelu_cases(), ids=[c[0] for c in cases.filtered_lrelu_cases()])
def test_filtered_lrelu_vs_golden(golden, name, kw):
    _, _, fl = _ops()
    x, fu, fd, b = cases.filtered_lrelu_inputs(name, kw)
    g = golden('filtered_lrelu')
    xt = cu(x).requires_grad_(True); bt = cu(b).requires_grad_(True)
    y = fl.filtered_lrelu(xt, cu(fu), cu(fd), bt, up=kw['up'], down=kw['down'], padding=kw['padding'], gain=kw['gain'], slope=kw['slope'], clamp=kw['clamp'])
    assert maxrel(y.detach().cpu().numpy(), g[name + '/y']) < TOL
    dy = cu(cases.cotangent(y.shape, 13))
    gx, gb = torch.autograd.grad(y, [xt, bt], dy)
    assert maxrel(gx.cpu().numpy(), g[name + '/dx']) < 5e-5
    assert maxrel(gb.cpu().numpy(), g[name + '/db']) < 5e-5


def test_filtered_lrelu_separable_cases_run_the_fused_kernel(golden):
    """Separable filters must take the single-kernel route (return code 0), not the generic upfirdn2d -> act -> upfirdn2d composition; the fused kernel's
    sign codes equal the generic route's on the region the outputs use, forward and backward agree with it, and fp16 tensors are covered."""
    _, up, fl = _ops()
    fl._init()
    plugin = fl._plugin
    for name, kw in cases.filtered_lrelu_cases():
        x, fu, fd, b = cases.filtered_lrelu_inputs(name, kw)
        if fu.ndim != 1 and fu.shape != (1, 1) or fd.ndim != 1 and fd.shape != (1, 1):
            continue
        px0, px1, py0, py1 = kw['padding']
        for dt, tol in ((torch.float32, 1e-5), (torch.float16, 4e-3)):
            xt = cu(x).to(dt); bt = cu(b).to(dt)
            y, so, rc = plugin.filtered_lrelu(xt, cu(fu), cu(fd), bt, torch.empty(0, dtype=torch.uint8, device='cuda'), kw['up'], kw['down'], px0, px1, py0, py1,
                                              0, 0, float(kw['gain']), float(kw['slope']), float(kw['clamp']) if kw['clamp'] is not None else float('inf'), False, True)
            assert rc == 0, name
            assert maxrel(y.float().cpu().numpy(), golden('filtered_lrelu')[name + '/y']) < tol, (name, dt)
            if dt == torch.float32:     # sign codes vs the generic route's
                t = xt + bt.reshape(1, -1, 1, 1)
                t = up.upfirdn2d(t, cu(fu), up=kw['up'], padding=kw['padding'], gain=kw['up'] ** 2).contiguous()
                so_g = plugin.filtered_lrelu_act_(t, torch.empty(0, dtype=torch.uint8, device='cuda'), 0, 0, float(kw['gain']), float(kw['slope']),
                                                  float(kw['clamp']) if kw['clamp'] is not None else float('inf'), True)
                sh, sw4 = so.shape[2], so.shape[3]
                a, g_ = so.cpu().numpy(), so_g.cpu().numpy()[:, :, :sh, :sw4]
                # compare decoded codes over the active width (the last byte of a row may hold codes of columns the outputs never use)
                fdt = fd.shape[-1] - 1
                swa = y.shape[3] * kw['down'] - (kw['down'] - 1) + fdt
                dec = lambda m: np.stack([(m >> (2 * k)) & 3 for k in range(4)], axis=-1).reshape(m.shape[0], m.shape[1], m.shape[2], -1)[..., :swa]
                mism = (dec(a) != dec(g_)).mean()
                assert mism < 1e-3, (name, mism)      # codes differ only where the two summation orders straddle 0 / the clamp


@pytest.mark.parametrize('kw', [dict(up=1, down=1, padding=[1, 1, 1, 1], gain=4), dict(up=2, down=1, padding=[2, 1, 2, 1], gain=4),
                                dict(up=1, down=2, padding=[1, 1, 1, 1], gain=1), dict(up=1, down=1, padding=[2, 2, 2, 2], gain=1),
                                dict(up=2, down=1, padding=[2, 1, 2, 1], gain=4, flip_filter=True)])
@pytest.mark.parametrize('shape', [(2, 8, 13, 11), (1, 96, 16, 16), (3, 128, 9, 33), (1, 1024, 4, 4)])
def test_upfirdn2d_register_tiled_channels_last_kernel(kw, shape):
    """The C-minor 4x4 kernel (4 output pixels per thread) must reproduce the W-minor kernel bit for bit (same tap order) and the oracle."""
    _, up, _ = _ops()
    rs = np.random.RandomState(sum(shape))
    x = rs.standard_normal(shape).astype(np.float32)
    f = cases._f2d(cases.F1331)
    y_nchw = up.upfirdn2d(cu(x), cu(f), **kw)
    y_cl = up.upfirdn2d(cu(x).contiguous(memory_format=torch.channels_last), cu(f), **kw)
    assert y_cl.shape == y_nchw.shape
    assert torch.equal(y_cl.contiguous(), y_nchw)
    yo = R.upfirdn2d(x, f, **kw)
    assert maxrel(y_nchw.cpu().numpy(), yo) < TOL


@pytest.mark.parametrize('shape,pad,flip', [((2, 64, 37, 41), (1, 1, 1, 1), False), ((3, 32, 16, 16), (2, 2, 2, 2), True), ((1, 128, 65, 9), (1, 2, 0, 3), False)])
def test_fir4_tma_kernel_matches_generic_upfirdn2d_bit_for_bit(shape, pad, flip):
    """gp3d_fir4_nhwc (TMA-staged window, csrc/fir_tma.cu) against the register-tiled / generic kernels behind gp3d_upfirdn2d on the same
    channel-minor float32 tensors, plus its two fused output forms (demodulation epilogue, bf16 operand pair)."""
    import ctypes
    import importlib
    _lib = importlib.import_module('3dgp_b200._lib')
    up = importlib.import_module('3dgp_b200.torch_utils.ops.upfirdn2d')
    tcm = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
    L = _lib.lib()
    torch.manual_seed(11)
    N, C, H, W = shape
    x = torch.randn(N, H, W, C, device='cuda')
    f = up.setup_filter([1, 3, 3, 1], device='cuda') * torch.linspace(0.5, 1.5, 16, device='cuda').view(4, 4)    # not symmetric: flip matters
    px0, px1, py0, py1 = pad
    oh, ow = H + py0 + py1 - 3, W + px0 + px1 - 3
    y = torch.empty(N, oh, ow, C, device='cuda')
    s = _lib.stream_ptr()
    _lib.check(L.gp3d_fir4_nhwc(x.data_ptr(), f.data_ptr(), int(flip), 4.0, N, H, W, C, px0, px1, py0, py1, y.data_ptr(), None, None, None, s), 'fir4')
    # reference: the same op on a tensor whose strides keep it off the TMA route (channel count padded in memory)
    xs = torch.zeros(N, H, W, C + 4, device='cuda')[..., :C]
    xs.copy_(x)
    ref = up._plugin.upfirdn2d(xs.permute(0, 3, 1, 2), f, 1, 1, 1, 1, px0, px1, py0, py1, flip, 4.0).permute(0, 2, 3, 1)
    assert ref.shape == y.shape and torch.equal(y, ref.contiguous())
    # routed automatically for dense tensors
    y2 = up._plugin.upfirdn2d(x.permute(0, 3, 1, 2), f, 1, 1, 1, 1, px0, px1, py0, py1, flip, 4.0).permute(0, 2, 3, 1)
    assert torch.equal(y2.contiguous(), y)
    # fused epilogue == demod_act on the filtered tensor
    d = torch.rand(N, C, device='cuda') + 0.5; nz = torch.randn(N, oh, ow, device='cuda') * 0.2; b = torch.randn(C, device='cuda') * 0.1
    y_ref = torch.empty_like(y); y_f = torch.empty_like(y)
    _lib.check(L.gp3d_demod_act(y.data_ptr(), d.data_ptr(), nz.data_ptr(), 1, b.data_ptr(), y_ref.data_ptr(), 0, N, C, oh * ow, 1, 3, 0.2, 1.4142135, -1.0, s), 'demod_act')
    epi = _lib.ConvEpilogue(d.data_ptr(), nz.data_ptr(), b.data_ptr(), 1, 3, 0.2, 1.4142135)
    _lib.check(L.gp3d_fir4_nhwc(x.data_ptr(), f.data_ptr(), int(flip), 4.0, N, H, W, C, px0, px1, py0, py1, y_f.data_ptr(), None, None, ctypes.byref(epi), s), 'fir4 epi')
    assert torch.equal(y_f, y_ref)
    # bf16 pair == split of the filtered tensor
    hi = torch.empty(N, oh, ow, C, dtype=torch.bfloat16, device='cuda'); lo = torch.empty_like(hi)
    _lib.check(L.gp3d_fir4_nhwc(x.data_ptr(), f.data_ptr(), int(flip), 4.0, N, H, W, C, px0, px1, py0, py1, None, hi.data_ptr(), lo.data_ptr(), None, s), 'fir4 split')
    rh, rl = tcm.split_bf16(y)
    assert torch.equal(hi, rh) and torch.equal(lo, rl)


@pytest.mark.parametrize('shape,pad,flip', [((2, 96, 16, 16), (2, 1, 2, 1), False), ((1, 32, 37, 21), (1, 2, 2, 1), True), ((3, 64, 5, 70), (2, 1, 1, 2), False)])
def test_fir4_up2_tma_kernel_matches_register_tiled_kernel_bit_for_bit(shape, pad, flip):
    """up = 2 form (upsample2d of the skip image): dense channel-minor float32 tensors take the TMA-staged kernel, tensors with a padded
    channel stride the register-tiled one; both must agree exactly, and with the oracle's zero-stuffing definition."""
    import importlib
    up = importlib.import_module('3dgp_b200.torch_utils.ops.upfirdn2d')
    from oracle import restated as R
    torch.manual_seed(12)
    N, C, H, W = shape
    x = torch.randn(N, H, W, C, device='cuda')
    f = up.setup_filter([1, 3, 3, 1], device='cuda') * torch.linspace(0.5, 1.5, 16, device='cuda').view(4, 4)
    px0, px1, py0, py1 = pad
    y = up._plugin.upfirdn2d(x.permute(0, 3, 1, 2), f, 2, 2, 1, 1, px0, px1, py0, py1, flip, 4.0)
    xs = torch.zeros(N, H, W, C + 4, device='cuda')[..., :C]
    xs.copy_(x)
    ref = up._plugin.upfirdn2d(xs.permute(0, 3, 1, 2), f, 2, 2, 1, 1, px0, px1, py0, py1, flip, 4.0)
    assert y.shape == ref.shape == (N, C, 2 * H + py0 + py1 - 3, 2 * W + px0 + px1 - 3)
    assert torch.equal(y.contiguous(), ref.contiguous())
    o = R.upfirdn2d(x.permute(0, 3, 1, 2).cpu().numpy(), f.cpu().numpy(), up=2, padding=[px0, px1, py0, py1], flip_filter=flip, gain=4)
    assert np.abs(y.cpu().numpy() - o).max() <= 2e-5 * np.abs(o).max()


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('shape', [(2, 3, 40, 48), (1, 5, 17, 136), (3, 2, 70, 264)])
@pytest.mark.parametrize('mode,flip', [('up', False), ('down', False), ('up', True), ('down', True)])
def test_nchw_tma_resampling_kernels_bit_exact_on_integers(mode, flip, shape, dtype):
    """upsample2d / downsample2d of dense NCHW planes (W % 8 == 0: the TMA-staged register-blocked kernels of csrc/fir_nchw_tma.cu) against the oracle's
    explicit-index upfirdn2d on integer-valued data with an asymmetric integer filter: every sum is exact in fp32 / fp16, so equality is bit-exact --
    index arithmetic, polyphase tap selection, convolution-vs-correlation and ragged tile edges included."""
    from oracle import restated as R
    _, up, _ = _ops()
    rs = np.random.RandomState(shape[2] * 7 + shape[3])
    x = rs.randint(-3, 4, size=shape).astype(np.float32)
    f = np.outer([1, 2, 3, 1], [2, 1, 3, 1]).astype(np.float32)             # asymmetric: flips and transposes are visible
    xt, ft = cu(x).to(dtype), cu(f)
    if mode == 'up':
        y = up.upsample2d(xt, ft, flip_filter=flip)
        ref = R.upfirdn2d(x, f, up=2, padding=[2, 1, 2, 1], flip_filter=flip, gain=4)
    else:
        y = up.downsample2d(xt, ft, flip_filter=flip)
        ref = R.upfirdn2d(x, f, down=2, padding=[1, 1, 1, 1], flip_filter=flip, gain=1)
    assert y.dtype == dtype and tuple(y.shape) == ref.shape
    assert np.array_equal(y.float().cpu().numpy(), ref)
