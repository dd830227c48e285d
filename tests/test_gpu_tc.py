"""GPU suite for the tcgen05 / TMEM contraction engine (csrc/conv_tc.cu, wgrad_tc.cu)."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


def _tc():
    return importlib.import_module('3dgp_b200.torch_utils.ops.tc')


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


@pytest.mark.parametrize('N,Cin,Cout,H,ks', [(2, 128, 128, 32, 3), (4, 64, 256, 16, 3), (2, 128, 96, 32, 1), (8, 1024 // 4, 1024 // 4, 4, 3)])
def test_conv2d_gradfix_tc_path_matches_aten_fp32(N, Cin, Cout, H, ks):
    """conv2d_gradfix routes eligible fp32 convs (forward and input-gradient) to the bf16x3 tcgen05 kernel: fp32-grade results."""
    cg = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix')
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device='cuda').manual_seed(Cin + Cout + H)
    x = torch.randn(N, Cin, H, H, device='cuda', generator=g).requires_grad_(True)
    w = (torch.randn(Cout, Cin, ks, ks, device='cuda', generator=g) / (ks * Cin ** 0.5)).requires_grad_(True)
    dy = torch.randn(N, Cout, H, H, device='cuda', generator=g)
    cg.tc_enabled = True
    before = dict(cg.tc_stats)
    y = cg.conv2d(x, w, padding=ks // 2)
    gx, gw = torch.autograd.grad(y, [x, w], dy)
    used_tc = cg.tc_stats['tc'] - before['tc']
    cg.tc_enabled = False
    y0 = cg.conv2d(x, w, padding=ks // 2)
    gx0, gw0 = torch.autograd.grad(y0, [x, w], dy)
    cg.tc_enabled = True
    assert used_tc >= 1
    rel = lambda a, b: (a - b).abs().max().item() / b.abs().max().item()
    assert rel(y, y0) < 1e-4 and rel(gx, gx0) < 1e-4 and rel(gw, gw0) < 1e-4, (rel(y, y0), rel(gx, gx0), rel(gw, gw0))


def test_conv2d_gradfix_tc_double_backward_r1_style():
    """R1-style second order through the tensor-core path: d/dw of |d logits / d x|^2 with weight gradients disabled in the inner pass."""
    cg = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix')
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device='cuda').manual_seed(5)
    w = (torch.randn(128, 64, 3, 3, device='cuda', generator=g) / 24).requires_grad_(True)
    x = torch.randn(2, 64, 16, 16, device='cuda', generator=g).requires_grad_(True)

    def r1(enabled):
        cg.tc_enabled = enabled
        y = cg.conv2d(x, w, padding=1)
        with cg.no_weight_gradients():
            gx = torch.autograd.grad(y.tanh().sum(), x, create_graph=True)[0]
        (gw,) = torch.autograd.grad(gx.square().sum(), w)
        return gw
    n_aten = cg.tc_stats['aten']
    a = r1(True)
    assert cg.tc_stats['aten'] == n_aten          # every conv of the second-order graph (incl. the weight gradient of the input-gradient op) ran on tcgen05
    b = r1(False)
    cg.tc_enabled = True
    assert (a - b).abs().max().item() / b.abs().max().item() < 2e-4


def test_conv2d_gradfix_fp16_uses_single_term_bf16():
    cg = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix')
    g = torch.Generator(device='cuda').manual_seed(9)
    x = torch.randn(2, 128, 32, 32, device='cuda', generator=g).half()
    w = (torch.randn(128, 128, 3, 3, device='cuda', generator=g) / 34).half()
    y = cg.conv2d(x, w, padding=1)
    ref = torch.nn.functional.conv2d(x.float(), w.float(), padding=1)
    assert y.dtype == torch.float16
    assert (y.float() - ref).abs().max().item() / ref.abs().max().item() < 2e-2


@pytest.mark.parametrize('N,Cin,Cout,H', [(2, 64, 128, 16), (3, 128, 128, 4), (1, 256, 128, 64)])
def test_transposed_conv_stride2_polyphase_and_its_gradients(N, Cin, Cout, H):
    """G's up-sampling conv0 (conv2d_resample.py:112-126): stride-2 transposed conv as four polyphase tap-convs; its input
    gradient is the strided-gather conv.  Both against ATen fp32."""
    cg = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix')
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device='cuda').manual_seed(N + Cin + H)
    x = torch.randn(N, Cin, H, H, device='cuda', generator=g).requires_grad_(True)
    w = (torch.randn(Cin, Cout, 3, 3, device='cuda', generator=g) / (3 * Cin ** 0.5)).requires_grad_(True)
    dy = torch.randn(N, Cout, 2 * H + 1, 2 * H + 1, device='cuda', generator=g)
    outs = []
    for en in (True, False):
        cg.tc_enabled = en
        n0 = cg.tc_stats['tc']
        y = cg.conv_transpose2d(x, w, stride=2, padding=0)
        gx, gw = torch.autograd.grad(y, [x, w], dy)
        outs.append((y, gx, gw, cg.tc_stats['tc'] - n0))
    cg.tc_enabled = True
    assert outs[0][3] >= 2 and outs[1][3] == 0
    rel = lambda a, b: (a - b).abs().max().item() / b.abs().max().item()
    assert tuple(outs[0][0].shape) == (N, Cout, 2 * H + 1, 2 * H + 1)
    for i in range(3):
        assert rel(outs[0][i], outs[1][i]) < 1e-4, (i, rel(outs[0][i], outs[1][i]))


@pytest.mark.parametrize('N,Cin,Cout,H,dtype', [(2, 128, 128, 36, torch.float32), (4, 64, 256, 12, torch.float32), (2, 128, 128, 20, torch.float16), (3, 64, 512, 12, torch.float16)])
def test_strided_conv_and_its_transposed_gradient(N, Cin, Cout, H, dtype):
    """D's down-sampling conv1 (FIR-padded input, stride-2 conv, conv2d_resample.py:106-109) and its input gradient (polyphase with output padding)."""
    cg = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix')
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device='cuda').manual_seed(N + Cin + H)
    x = torch.randn(N, Cin, H, H, device='cuda', generator=g).to(dtype).requires_grad_(True)
    w = (torch.randn(Cout, Cin, 3, 3, device='cuda', generator=g) / (3 * Cin ** 0.5)).to(dtype).requires_grad_(True)
    outs = []
    for en in (True, False):
        cg.tc_enabled = en
        y = cg.conv2d(x, w, stride=2)
        dy = torch.ones_like(y) * 0.5
        dy[..., ::2, :] *= -1
        gx, gw = torch.autograd.grad(y, [x, w], dy)
        outs.append((y.float(), gx.float(), gw.float()))
    cg.tc_enabled = True
    tol = 1e-4 if dtype == torch.float32 else 3e-2
    rel = lambda a, b: (a - b).abs().max().item() / b.abs().max().item()
    for i in range(3):
        assert outs[0][i].shape == outs[1][i].shape
        assert rel(outs[0][i], outs[1][i]) < tol, (i, rel(outs[0][i], outs[1][i]))


@pytest.mark.parametrize('up,noise_mode', [(1, 'random'), (2, 'const'), (2, 'random'), (1, 'none')])
def test_fused_modconv_layer_matches_unfused_training_path(up, noise_mode):
    """ops/modconv.py (one autograd node: modulate+split -> tcgen05 conv -> FIR -> demod+noise+bias+lrelu, fused backward) against the
    reference-shaped composition x*styles -> conv2d_resample -> fma -> bias_act, same weights: outputs and every gradient."""
    sg = importlib.import_module('3dgp_b200.training.networks_stylegan2')
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(3)
    cin, cout, res, B, wd = (128, 128, 32, 3, 64) if up == 1 else (256, 128, 32, 2, 64)
    layer = sg.SynthesisLayer(cin, cout, w_dim=wd, resolution=res, up=up, conv_clamp=None).cuda()
    with torch.no_grad():
        layer.noise_strength.fill_(0.3); layer.bias.normal_(0, 0.2)
    x = torch.randn(B, cin, res // up, res // up, device='cuda', requires_grad=True)
    w = torch.randn(B, wd, device='cuda', requires_grad=True)
    nz = torch.randn(B, 1, res, res, device='cuda')
    dy = torch.randn(B, cout, res, res, device='cuda')
    res_ = []
    for fused in (True, False):
        sg.fused_layer_enabled = fused
        y = layer(x, w, noise_mode=noise_mode, fused_modconv=False, noise_in=nz)
        params = [x, w, layer.weight, layer.bias, layer.affine.weight, layer.affine.bias] + ([layer.noise_strength] if noise_mode != 'none' else [])
        gs = torch.autograd.grad(y, params, dy)
        res_.append((y, gs))
    sg.fused_layer_enabled = True
    rel = lambda a, b: (a - b).abs().max().item() / max(b.abs().max().item(), 1e-20)
    assert rel(res_[0][0], res_[1][0]) < 1e-4
    for a, b in zip(res_[0][1], res_[1][1]):
        assert a.shape == b.shape and rel(a, b) < 3e-4, (a.shape, rel(a, b))


def test_fused_torgb_matches_unfused():
    sg = importlib.import_module('3dgp_b200.training.networks_stylegan2')
    torch.manual_seed(4)
    layer = sg.ToRGBLayer(128, 96, w_dim=64).cuda()
    x = torch.randn(2, 128, 32, 32, device='cuda', requires_grad=True)
    w = torch.randn(2, 64, device='cuda', requires_grad=True)
    dy = torch.randn(2, 96, 32, 32, device='cuda')
    out = []
    for fused in (True, False):
        sg.fused_layer_enabled = fused
        y = layer(x, w, fused_modconv=False)
        out.append((y, torch.autograd.grad(y, [x, w, layer.weight, layer.bias], dy)))
    sg.fused_layer_enabled = True
    rel = lambda a, b: (a - b).abs().max().item() / max(b.abs().max().item(), 1e-20)
    assert rel(out[0][0], out[1][0]) < 1e-4
    for a, b in zip(out[0][1], out[1][1]):
        assert rel(a, b) < 3e-4


def test_conv5x5_depth_adaptor_shape_forward_and_input_gradient():
    """5x5 'same' convs of the depth adaptor (networks_depth_adaptor.py:31-33: 64 -> 64 channels at the patch resolution), forward,
    input gradient and weight gradient on the tcgen05 path.  Checked against a float64 CPU convolution: cuDNN's own fp32 5x5 kernel
    for this shape is only accurate to ~1e-2 on B200 (measured, tools/debug_conv5.py), so it cannot serve as the comparator."""
    cg = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix')
    g = torch.Generator(device='cuda').manual_seed(11)
    x = torch.randn(4, 64, 64, 64, device='cuda', generator=g).requires_grad_(True)
    w = (torch.randn(64, 64, 5, 5, device='cuda', generator=g) / 40).requires_grad_(True)
    dy = torch.randn(4, 64, 64, 64, device='cuda', generator=g)
    cg.tc_enabled = True
    n0 = cg.tc_stats['tc']
    y = cg.conv2d(x, w, padding=2)
    gx, gw = torch.autograd.grad(y, [x, w], dy)
    assert cg.tc_stats['tc'] - n0 >= 3          # forward, input gradient and weight gradient all on the tensor-core path
    xd = x.detach().double().cpu().requires_grad_(True); wd = w.detach().double().cpu().requires_grad_(True)
    yd = torch.nn.functional.conv2d(xd, wd, padding=2)
    gxd, gwd = torch.autograd.grad(yd, [xd, wd], dy.double().cpu())
    rel = lambda a, b: ((a.double().cpu() - b).abs().max() / b.abs().max()).item()
    assert rel(y, yd.detach()) < 5e-5 and rel(gx, gxd) < 5e-5 and rel(gw, gwd) < 5e-5, (rel(y, yd.detach()), rel(gx, gxd), rel(gw, gwd))


def test_split_bf16_zero_pads_the_channel_tail():
    tc = _tc()
    torch.manual_seed(0)
    x = torch.randn(3, 5, 7, 96, device='cuda')
    s = torch.rand(3, 96, device='cuda') + 0.5
    hi, lo = tc.split_bf16(x, styles=s, pad_to=128)
    assert hi.shape == (3, 5, 7, 128) and lo.shape == hi.shape
    assert (hi[..., 96:] == 0).all() and (lo[..., 96:] == 0).all()
    ref = x * s[:, None, None, :]
    got = hi[..., :96].float() + lo[..., :96].float()
    assert (got - ref).abs().max().item() <= ref.abs().max().item() * 2 ** -15
    h2, l2 = tc.split_bf16(x, styles=s)
    assert torch.equal(h2, hi[..., :96]) and torch.equal(l2, lo[..., :96])


def test_weight_operand_cache_follows_in_place_updates():
    tc = _tc()
    w = torch.nn.Parameter(torch.randn(128, 64, 3, 3, device='cuda'))
    a = tc.weight_operands(w, 'fwd', lambda t: t.permute(0, 2, 3, 1))
    b = tc.weight_operands(w, 'fwd', lambda t: t.permute(0, 2, 3, 1))
    assert a[0] is b[0] and a[1] is b[1]                                   # one split per optimiser step
    with torch.no_grad():
        w.mul_(2.0)                                                         # what an optimiser does
    c = tc.weight_operands(w, 'fwd', lambda t: t.permute(0, 2, 3, 1))
    assert c[0] is not a[0]
    assert torch.equal(c[0].float(), a[0].float() * 2)
    t = torch.randn(128, 64, 3, 3, device='cuda')                           # temporaries are never cached
    assert tc.weight_operands(t, 'fwd', lambda u: u.permute(0, 2, 3, 1))[0] is not tc.weight_operands(t, 'fwd', lambda u: u.permute(0, 2, 3, 1))[0]


@pytest.mark.parametrize('res,up', [(4, 1), (8, 2)])
def test_fused_modconv_layer_at_the_coarsest_resolutions(res, up):
    """b4.conv1 (4x4) and b8.conv0 (4x4 -> 8x8) run the same fused node as the large layers."""
    sg = importlib.import_module('3dgp_b200.training.networks_stylegan2')
    torch.manual_seed(5)
    layer = sg.SynthesisLayer(128, 128, w_dim=64, resolution=res, up=up, conv_clamp=None).cuda()
    with torch.no_grad():
        layer.noise_strength.fill_(0.2)
    B = 5
    x = torch.randn(B, 128, res // up, res // up, device='cuda', requires_grad=True)
    w = torch.randn(B, 64, device='cuda', requires_grad=True)
    nz = torch.randn(B, 1, res, res, device='cuda')
    dy = torch.randn(B, 128, res, res, device='cuda')
    out = []
    for fused in (True, False):
        sg.fused_layer_enabled = fused
        y = layer(x, w, noise_mode='random', fused_modconv=False, noise_in=nz)
        out.append((y, torch.autograd.grad(y, [x, w, layer.weight, layer.bias, layer.noise_strength], dy)))
    sg.fused_layer_enabled = True
    rel = lambda a, b: (a - b).abs().max().item() / max(b.abs().max().item(), 1e-20)
    assert rel(out[0][0], out[1][0]) < 1e-4
    for a, b in zip(out[0][1], out[1][1]):
        assert rel(a, b) < 3e-4, rel(a, b)


@pytest.mark.parametrize('Cout,act,with_d,nps', [(128, 3, True, 1), (96, 1, False, 0), (256, 3, True, 0)])
def test_conv_fused_epilogue_matches_conv_then_demod_act(Cout, act, with_d, nps):
    """gp3d_conv2d_nhwc_bf16x3_act == gp3d_conv2d_nhwc_bf16x3 followed by gp3d_demod_act (bit-for-bit: same fma order)."""
    import ctypes
    tc = _tc()
    _lib = importlib.import_module('3dgp_b200._lib')
    L = _lib.lib()
    torch.manual_seed(7)
    N, H, W, Cin, k = 3, 20, 24, 64, 3
    x = torch.randn(N, H, W, Cin, device='cuda'); w = torch.randn(Cout, k, k, Cin, device='cuda') / (Cin * 9) ** 0.5
    xh, xl = tc.split_bf16(x); wh, wl = tc.split_bf16(w)
    d = (torch.rand(N, Cout, device='cuda') + 0.5) if with_d else None
    nz = torch.randn(N if nps else 1, H, W, device='cuda') * 0.3
    b = torch.randn(Cout, device='cuda') * 0.1
    c = torch.empty(N, H, W, Cout, device='cuda'); y_ref = torch.empty_like(c); y = torch.empty_like(c)
    s = _lib.stream_ptr()
    _lib.check(L.gp3d_conv2d_nhwc_bf16x3(xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr(), c.data_ptr(), N, H, W, Cin, Cout, k, 0, s), 'conv')
    _lib.check(L.gp3d_demod_act(c.data_ptr(), _lib.ptr(d), nz.data_ptr(), nps, b.data_ptr(), y_ref.data_ptr(), 0, N, Cout, H * W, 1, act, 0.2, 1.4142135, -1.0, s), 'demod_act')
    epi = _lib.ConvEpilogue(_lib.ptr(d), nz.data_ptr(), b.data_ptr(), nps, act, 0.2, 1.4142135)
    _lib.check(L.gp3d_conv2d_nhwc_bf16x3_act(xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr(), y.data_ptr(), N, H, W, Cin, Cout, k, ctypes.byref(epi), s), 'conv_act')
    assert torch.equal(y, y_ref)


def test_demod_act_bwd_split_outputs_equal_split_of_float_outputs():
    tc = _tc()
    _lib = importlib.import_module('3dgp_b200._lib')
    L = _lib.lib()
    torch.manual_seed(8)
    N, HW, C, Cp = 2, 37, 96, 128
    dy = torch.randn(N, HW, C, device='cuda'); y = torch.randn(N, HW, C, device='cuda'); d = torch.rand(N, C, device='cuda') + 0.5
    b = torch.randn(C, device='cuda') * 0.1
    dc = torch.empty_like(dy)
    outs = []
    for split in (False, True):
        g_d = torch.zeros_like(d); g_b = torch.zeros(C, device='cuda')
        hi = torch.full([N, HW, Cp], 7.0, dtype=torch.bfloat16, device='cuda'); lo = torch.full_like(hi, 7.0)
        rc = L.gp3d_demod_act_bwd_split(dy.data_ptr(), y.data_ptr(), d.data_ptr(), None, None, 0, b.data_ptr(), None if split else dc.data_ptr(),
                                        hi.data_ptr() if split else None, lo.data_ptr() if split else None, Cp, g_d.data_ptr(), g_b.data_ptr(), None,
                                        N, HW, C, 3, 0.2, 1.4142135, _lib.stream_ptr())
        _lib.check(rc, 'demod_act_bwd_split')
        outs.append((g_d, g_b, hi, lo))
    rh, rl = tc.split_bf16(dc, pad_to=Cp)
    assert torch.equal(outs[1][2], rh) and torch.equal(outs[1][3], rl)
    assert torch.equal(outs[0][0], outs[1][0]) or torch.allclose(outs[0][0], outs[1][0], rtol=1e-5, atol=1e-5)
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('cl', [True, False])
def test_channel_scale_fused_backward_matches_autograd(cl):
    """Hyper-modulation x * s[n, c] of Conv2dLayer (layers.py:231-232): one-kernel backward against plain autograd."""
    layers = importlib.import_module('3dgp_b200.training.layers')
    torch.manual_seed(9)
    x = torch.randn(3, 64, 10, 12, device='cuda')
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    s = (1 + torch.randn(3, 64, device='cuda').tanh()).requires_grad_(True)
    dy = torch.randn(3, 64, 10, 12, device='cuda')
    y = layers._ChannelScale.apply(x, s)
    gx, gs = torch.autograd.grad(y, [x, s], dy)
    yr = x * s.unsqueeze(2).unsqueeze(3)
    rx, rs = torch.autograd.grad(yr, [x, s], dy)
    assert torch.equal(y, yr) and torch.allclose(gx, rx, rtol=1e-6, atol=1e-6) and torch.allclose(gs, rs, rtol=1e-4, atol=1e-4)


def test_conv_and_fir_at_baseline_size_exact_properties():
    """BASELINE layer shape (b512.conv1: 128 -> 128 channels at 512^2; the FIR of b512.conv0 on 513^2): size-independent exact properties.
    Scaling the input by 2 scales every bf16 (hi, lo) operand by exactly 2, so the convolution must double bit-for-bit; the fused
    epilogue must equal conv followed by demod_act; the TMA-staged FIR must equal the register-tiled kernel (reached through a padded stride)."""
    import ctypes
    tc = _tc()
    _lib = importlib.import_module('3dgp_b200._lib')
    up = importlib.import_module('3dgp_b200.torch_utils.ops.upfirdn2d')
    L = _lib.lib()
    s = _lib.stream_ptr()
    torch.manual_seed(21)
    N, H, C = 2, 512, 128
    x = torch.randn(N, H, H, C, device='cuda'); w = torch.randn(C, 3, 3, C, device='cuda') / 34
    wh, wl = tc.split_bf16(w)
    ys = []
    for scale in (1.0, 2.0):
        xh, xl = tc.split_bf16(x * scale)
        y = torch.empty(N, H, H, C, device='cuda')
        _lib.check(L.gp3d_conv2d_nhwc_bf16x3(xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr(), y.data_ptr(), N, H, H, C, C, 3, 0, s), 'conv')
        ys.append(y)
    assert torch.isfinite(ys[0]).all() and torch.equal(ys[1], ys[0] * 2)
    d = torch.rand(N, C, device='cuda') + 0.5; nz = torch.randn(H, H, device='cuda') * 0.1; b = torch.randn(C, device='cuda') * 0.1
    ref = torch.empty_like(ys[0]); got = torch.empty_like(ys[0])
    _lib.check(L.gp3d_demod_act(ys[0].data_ptr(), d.data_ptr(), nz.data_ptr(), 0, b.data_ptr(), ref.data_ptr(), 0, N, C, H * H, 1, 3, 0.2, 1.4142135, -1.0, s), 'demod')
    xh, xl = tc.split_bf16(x)
    epi = _lib.ConvEpilogue(d.data_ptr(), nz.data_ptr(), b.data_ptr(), 0, 3, 0.2, 1.4142135)
    _lib.check(L.gp3d_conv2d_nhwc_bf16x3_act(xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr(), got.data_ptr(), N, H, H, C, C, 3, ctypes.byref(epi), s), 'conv_act')
    assert torch.equal(got, ref)
    del ys, ref, got, xh, xl
    f = up.setup_filter([1, 3, 3, 1], device='cuda')
    c1 = torch.randn(N, 513, 513, C, device='cuda')
    a = up._plugin.upfirdn2d(c1.permute(0, 3, 1, 2), f, 1, 1, 1, 1, 1, 1, 1, 1, False, 4.0)
    c1s = torch.zeros(N, 513, 513, C + 4, device='cuda')[..., :C]
    c1s.copy_(c1)
    b2 = up._plugin.upfirdn2d(c1s.permute(0, 3, 1, 2), f, 1, 1, 1, 1, 1, 1, 1, 1, False, 4.0)
    assert a.shape == (N, C, 512, 512) and torch.equal(a.contiguous(), b2.contiguous())


@pytest.mark.parametrize('terms', [3, 1])
@pytest.mark.parametrize('k,act,hyper,bias,clamp,gain', [(3, 'lrelu', True, True, 256, 1.0), (1, 'linear', False, False, None, 0.7071), (3, 'lrelu', False, True, 0.5, 1.0)])
def test_fused_conv2d_layer_matches_unfused(terms, k, act, hyper, bias, clamp, gain):
    """Conv2dLayer (layers.py:228-241) through ops/modconv.py::_ConvBiasAct against the reference-shaped composition
    hyper-mod * x -> conv2d_gradfix -> bias_act on the same weights: output and every gradient, both precisions."""
    layers = importlib.import_module('3dgp_b200.training.layers')
    cg = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix')
    torch.manual_seed(13)
    layer = layers.Conv2dLayer(128, 256, kernel_size=k, bias=bias, activation=act, conv_clamp=clamp, c_dim=32 if hyper else 0, hyper_mod=hyper).cuda()
    with torch.no_grad():
        if bias:
            layer.bias.normal_(0, 0.3)
    x = torch.randn(3, 128, 20, 24, device='cuda').contiguous(memory_format=torch.channels_last).requires_grad_(True)
    c = torch.randn(3, 32, device='cuda') if hyper else None
    dy = torch.randn(3, 256, 20, 24, device='cuda')
    params = [x] + [p for p in layer.parameters()]
    outs = []
    for fused in (True, False):
        with layers.first_order_only(fused), cg.tc_terms(terms):
            y = layer(x, c=c, gain=gain)
            gs = torch.autograd.grad(y, params, dy)
        outs.append((y, gs))
    rel = lambda a, b: (a - b).abs().max().item() / max(b.abs().max().item(), 1e-20)
    tol = 3e-4 if terms == 3 else 2e-3
    assert outs[0][0].shape == outs[1][0].shape and rel(outs[0][0], outs[1][0]) < tol
    for a, b in zip(outs[0][1], outs[1][1]):
        assert a.shape == b.shape and rel(a, b) < tol, (a.shape, rel(a, b))
    if clamp is not None and clamp < 1:
        assert (outs[0][0].abs() <= clamp * gain + 1e-6).all() and (outs[0][0].abs() >= clamp * gain - 1e-6).any()     # the clamp is active in this case
    y2 = layer(x, c=c, gain=gain)
    y2.add_(1.0)                                             # in-place update of the output (DiscriminatorBlock: y.add_(x)) must be legal


@pytest.mark.parametrize('N,Cin,Cout,H,k', [(2, 128, 256, 32, 3), (1, 256, 512, 16, 3), (4, 64, 256, 8, 1), (1, 64, 1024, 16, 3)])
def test_three_term_conv_256_wide_tiles_equal_128_wide_tiles(N, Cin, Cout, H, k):
    """The bf16x3 form with 256-wide output-channel tiles (two-stage ring of 96 KB stages) accumulates every output element over the same K order as the
    128-wide form: bit-identical results, and fp32-grade against float64."""
    tc = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
    _lib = importlib.import_module('3dgp_b200._lib')
    torch.manual_seed(Cout + H)
    x = torch.randn(N, Cin, H, H, device='cuda'); w = torch.randn(Cout, Cin, k, k, device='cuda') / (Cin * k * k) ** 0.5
    L = _lib.lib()
    old = L.gp3d_conv_set_wide3(1)
    try:
        y_wide = tc.conv2d_forward(x, w, 3)
        L.gp3d_conv_set_wide3(0)
        y_narrow = tc.conv2d_forward(x, w, 3)
    finally:
        L.gp3d_conv_set_wide3(old)
    assert torch.equal(y_wide, y_narrow)
    yd = torch.nn.functional.conv2d(x.double(), w.double(), padding=k // 2)
    assert ((y_wide.double() - yd).norm() / yd.norm()).item() < 2e-5


@pytest.mark.parametrize('terms', [3, 1])
@pytest.mark.parametrize('N,Cin,Cout,H,k', [(2, 256, 128, 32, 3), (1, 512, 256, 16, 3), (4, 256, 64, 8, 1)])
def test_weight_gradient_256_wide_cin_tiles(N, Cin, Cout, H, k, terms):
    """wgrad_kernel<*, 256> (Cin % 256 == 0): same products as the 128-wide form (split-K partition differs, so the sums agree to rounding) and
    fp32-grade (three-term) against float64."""
    tc = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
    L = importlib.import_module('3dgp_b200._lib').lib()
    torch.manual_seed(Cin + H + terms)
    x = torch.randn(N, Cin, H, H, device='cuda'); dy = torch.randn(N, Cout, H, H, device='cuda')
    old = L.gp3d_conv_set_wide3(1)
    try:
        gw_wide = tc.conv_wgrad(dy, x, k, 'conv', 1, k // 2, terms)
        L.gp3d_conv_set_wide3(0)
        gw_narrow = tc.conv_wgrad(dy, x, k, 'conv', 1, k // 2, terms)
    finally:
        L.gp3d_conv_set_wide3(old)
    w = torch.zeros(Cout, Cin, k, k, device='cuda', dtype=torch.float64, requires_grad=True)
    gwd = torch.autograd.grad(torch.nn.functional.conv2d(x.double(), w, padding=k // 2), w, dy.double())[0]
    rel = lambda a: ((a.double() - gwd).norm() / gwd.norm()).item()
    assert ((gw_wide - gw_narrow).norm() / gw_narrow.norm()).item() < 1e-5
    assert rel(gw_wide) < (2e-5 if terms == 3 else 6e-3)
