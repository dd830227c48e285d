"""GPU parity suite for the fused ray-march kernels (forward + backward) through the C ABI."""
import importlib

import numpy as np
import pytest
import torch

from oracle import cases, restated as R
from util import maxrel, l2rel

pytestmark = pytest.mark.gpu
TOL = 1e-3   # north_star: within 1e-3 relative fp32 of the reference on identical latent/camera/noise


def _rm():
    return importlib.import_module('3dgp_b200.torch_utils.ops.raymarch')


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _run(kw, inp, requires_grad=False, mlp_mode=0, planes_dtype=torch.float32):
    rm = _rm()
    t = {k: cu(v) for k, v in inp.items()}
    planes = t['planes'].to(planes_dtype)
    args = [planes, t['w1'], t['b1'], t['w2'], t['b2'], t['ray_o'], t['ray_d']]
    if requires_grad:
        args = [a.clone().requires_grad_(True) for a in args]
    out = rm.render_rays(*args, num_steps=kw['N'], ray_start=kw['ray_start'], ray_end=kw['ray_end'], box_size=2 * kw['box_half'],
                         u_coarse=t['u_coarse'], u_fine=t['u_fine'], sn_coarse=t.get('sn_coarse'), sn_fine=t.get('sn_fine'),
                         density_noise=kw.get('noise_std', 0.0), use_inf_depth=kw.get('use_inf_depth', True), last_back=kw.get('last_back', False),
                         white_back_end_idx=kw.get('white_back_end_idx', 0), clamp_mode=kw.get('clamp_mode', 'softplus'), mlp_mode=mlp_mode)
    return out, args


@pytest.mark.parametrize('name,kw', cases.render_cases(), ids=[c[0] for c in cases.render_cases()])
def test_render_forward_vs_golden(golden, name, kw):
    inp = cases.render_inputs(name, kw)
    (rgb, depth, wsum, tfin), _ = _run(kw, inp)
    g = golden('render')
    assert maxrel(rgb.cpu().numpy(), g[name + '/rgb']) < TOL
    assert maxrel(depth.squeeze(-1).cpu().numpy(), g[name + '/depth']) < TOL
    assert maxrel(wsum.squeeze(-1).cpu().numpy(), g[name + '/wsum']) < TOL
    assert maxrel(tfin.cpu().numpy(), g[name + '/tfinal']) < TOL
    # fp32 SIMT MLP path should in fact sit at re-association level
    assert maxrel(rgb.cpu().numpy(), g[name + '/rgb']) < 2e-5


@pytest.mark.parametrize('name,kw', cases.render_cases(), ids=[c[0] for c in cases.render_cases()])
def test_render_backward_vs_golden(golden, name, kw):
    inp = cases.render_inputs(name, kw)
    (rgb, depth, _, _), args = _run(kw, inp, requires_grad=True)
    g_rgb = cu(cases.cotangent(rgb.shape, 21)); g_dep = cu(cases.cotangent(depth.shape, 22))
    grads = torch.autograd.grad([rgb, depth], args, [g_rgb, g_dep])
    g = golden('render')
    names = ['g_planes', 'g_w1', 'g_b1', 'g_w2', 'g_b2', 'g_ray_o', 'g_ray_d']
    for nm, gr in zip(names, grads):
        gr = gr.contiguous() if nm != 'g_planes' else gr.permute(0, 1, 2, 3, 4).contiguous()
        if nm == 'g_planes' and (name + '/g_planes') not in g.files:
            probe = gr.flatten()[::97].cpu().numpy()
            assert l2rel(probe, g[name + '/g_planes_probe']) < TOL
            st = np.array([gr.double().sum().item(), gr.double().abs().sum().item(), gr.double().square().sum().item()])
            assert np.allclose(st[1:], g[name + '/g_planes_sum'][1:], rtol=1e-3)
            continue
        ref = g[name + '/' + nm]
        assert gr.shape == ref.shape, (nm, gr.shape, ref.shape)
        assert l2rel(gr.cpu().numpy(), ref) < TOL, nm
        assert maxrel(gr.cpu().numpy(), ref) < 5e-3, nm


def test_render_fresh_seed_vs_oracle_and_layouts():
    """Fresh inputs (not in the fixtures): CUDA vs CPU oracle; NCHW-strided planes (copied to channel-minor inside) and
    channels-last planes must agree bit-for-bit; fp16 plane storage within its own tolerance."""
    kw = dict(C=32, H=64, ray_start=0.75, ray_end=1.25, box_half=0.5, B=2, R=70, N=16, P=48)
    inp = cases.render_inputs('fresh_seed_1', kw)
    (rgb, depth, wsum, tfin), _ = _run(kw, inp)
    t = {k: torch.from_numpy(v) for k, v in inp.items()}
    o = R.render(t['planes'], t['w1'], t['b1'], t['w2'], t['b2'], t['ray_o'], t['ray_d'], t['u_coarse'], t['u_fine'], 0.75, 1.25, 0.5, 16)
    assert maxrel(rgb.cpu().numpy(), o[0]) < 2e-5 and maxrel(depth.squeeze(-1).cpu().numpy(), o[1]) < 2e-5
    rm = _rm()
    pl = cu(inp['planes'])
    plc = rm.planes_channel_minor(pl)
    assert plc.stride(2) == 1
    tt = {k: cu(v) for k, v in inp.items()}
    out2 = rm.render_rays(plc, tt['w1'], tt['b1'], tt['w2'], tt['b2'], tt['ray_o'], tt['ray_d'], num_steps=16, ray_start=0.75, ray_end=1.25,
                          box_size=1.0, u_coarse=tt['u_coarse'], u_fine=tt['u_fine'])
    assert torch.equal(out2[0], rgb)
    (rgb16, depth16, _, _), _ = _run(kw, inp, planes_dtype=torch.float16)
    assert maxrel(rgb16.cpu().numpy(), o[0]) < 2e-3 and maxrel(depth16.squeeze(-1).cpu().numpy(), o[1]) < 1e-3


def test_render_full_size_properties():
    """BASELINE config 3 geometry (512^2 x 32ch planes, 64x64 rays, 48+48 samples) on one image: invariants that do not
    need the oracle -- determinism, weights in [0,1], depth inside [ray_start, ray_end] when opacity saturates,
    in-kernel Philox mode reproducible per (seed, offset) and different across seeds."""
    rm = _rm()
    torch.manual_seed(0)
    B, P, Rr, N = 2, 512, 4096, 48
    planes = torch.randn([B, P, P, 96], device='cuda').permute(0, 3, 1, 2).view(B, 3, 32, P, P)
    w1 = torch.randn(64, 32, device='cuda'); b1 = torch.zeros(64, device='cuda'); w2 = torch.randn(4, 64, device='cuda'); b2 = torch.zeros(4, device='cuda')
    inp = cases.render_inputs('full_rays', dict(B=B, R=Rr, N=4, P=2, C=32, H=64))
    ro, rd = cu(inp['ray_o']), cu(inp['ray_d'])
    kw = dict(num_steps=N, ray_start=0.75, ray_end=1.25, box_size=1.0)
    a = rm.render_rays(planes, w1, b1, w2, b2, ro, rd, seed=7, **kw)
    b = rm.render_rays(planes, w1, b1, w2, b2, ro, rd, seed=7, **kw)
    c = rm.render_rays(planes, w1, b1, w2, b2, ro, rd, seed=8, **kw)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert not torch.equal(a[0], c[0])
    rgb, depth, wsum, tfin = a
    assert torch.isfinite(rgb).all() and torch.isfinite(depth).all()
    assert (wsum >= -1e-5).all() and (wsum <= 1 + 1e-4).all() and (tfin >= 0).all() and (tfin <= 1).all()
    sat = wsum.squeeze(-1) > 0.999
    assert sat.any()
    d = depth.squeeze(-1)[sat]
    assert (d > 0.75 - 1e-3).all() and (d < 1.25 + 1e-3).all()


@pytest.mark.parametrize('mode,tol', [(1, 2e-3), (2, 5e-5)])      # mode 1 = plain TF32 (10-bit operands): a throughput mode, not the default
@pytest.mark.parametrize('name,kw', cases.render_cases(), ids=[c[0] for c in cases.render_cases()])
def test_render_forward_tensor_core_mlp_vs_golden(golden, name, kw, mode, tol):
    """Second-generation forward kernel (raymarch_fwd2.cu): MLP on mma.sync TF32 (mode 1) / 3xTF32 (mode 2), parallel per-ray phases."""
    inp = cases.render_inputs(name, kw)
    (rgb, depth, wsum, tfin), _ = _run(kw, inp, mlp_mode=mode)
    g = golden('render')
    assert maxrel(rgb.cpu().numpy(), g[name + '/rgb']) < tol
    assert maxrel(depth.squeeze(-1).cpu().numpy(), g[name + '/depth']) < tol
    assert maxrel(wsum.squeeze(-1).cpu().numpy(), g[name + '/wsum']) < tol
    assert maxrel(tfin.cpu().numpy(), g[name + '/tfinal']) < tol


@pytest.mark.parametrize('mode,tol', [(1, 6e-3), (2, TOL)])
@pytest.mark.parametrize('name,kw', cases.render_cases(), ids=[c[0] for c in cases.render_cases()])
def test_render_backward_tensor_core_vs_golden(golden, name, kw, mode, tol):
    """Second-generation backward kernel (raymarch_bwd2.cu): recompute + every MLP contraction (layer 1, dW2, dW1, d feature) on
    mma.sync TF32 (mode 1) / 3xTF32 (mode 2), parallel compositing backward; all seven gradients against the reference's autograd."""
    inp = cases.render_inputs(name, kw)
    (rgb, depth, _, _), args = _run(kw, inp, requires_grad=True, mlp_mode=mode)
    g_rgb = cu(cases.cotangent(rgb.shape, 21)); g_dep = cu(cases.cotangent(depth.shape, 22))
    grads = torch.autograd.grad([rgb, depth], args, [g_rgb, g_dep], retain_graph=True)
    g = golden('render')
    names = ['g_planes', 'g_w1', 'g_b1', 'g_w2', 'g_b2', 'g_ray_o', 'g_ray_d']
    for nm, gr in zip(names, grads):
        gr = gr.contiguous()
        if nm == 'g_planes' and (name + '/g_planes') not in g.files:
            probe = gr.flatten()[::97].cpu().numpy()
            assert l2rel(probe, g[name + '/g_planes_probe']) < tol
            st = np.array([gr.double().sum().item(), gr.double().abs().sum().item(), gr.double().square().sum().item()])
            assert np.allclose(st[1:], g[name + '/g_planes_sum'][1:], rtol=max(tol, 1e-3))
            continue
        ref = g[name + '/' + nm]
        assert gr.shape == ref.shape, (nm, gr.shape, ref.shape)
        assert l2rel(gr.cpu().numpy(), ref) < tol, (nm, l2rel(gr.cpu().numpy(), ref))
        assert maxrel(gr.cpu().numpy(), ref) < 5 * tol, (nm, maxrel(gr.cpu().numpy(), ref))
    # the training configuration asks for no ray gradients (learn_camera_dist = false): same kernel without the tap dot-products
    g2 = torch.autograd.grad([rgb, depth], args[:5], [g_rgb, g_dep])
    for a, b in zip(g2, grads[:5]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6 * float(b.abs().max()))


def test_render_backward_generations_agree_at_full_size():
    """BASELINE config 3 geometry, one image: first-generation (fp32 SIMT) and second-generation (3xTF32) backward kernels on the
    same inputs in Philox mode (same seed => same samples)."""
    rm = _rm()
    torch.manual_seed(1)
    B, P, Rr, N = 1, 512, 4096, 48
    planes = (torch.randn([B, P, P, 96], device='cuda') * 0.5).permute(0, 3, 1, 2).view(B, 3, 32, P, P).requires_grad_(True)
    w1 = torch.randn(64, 32, device='cuda', requires_grad=True); b1 = (torch.randn(64, device='cuda') * 0.1).requires_grad_(True)
    w2 = torch.randn(4, 64, device='cuda', requires_grad=True); b2 = torch.zeros(4, device='cuda', requires_grad=True)
    inp = cases.render_inputs('full_rays', dict(B=B, R=Rr, N=4, P=2, C=32, H=64))
    ro, rd = cu(inp['ray_o']), cu(inp['ray_d'])
    kw = dict(num_steps=N, ray_start=0.75, ray_end=1.25, box_size=1.0, seed=11, density_noise=0.3)
    res = []
    for mode in (0, 2):
        rgb, depth, _, _ = rm.render_rays(planes, w1, b1, w2, b2, ro, rd, mlp_mode=mode, **kw)
        gr = torch.autograd.grad([rgb, depth], [planes, w1, b1, w2, b2], [torch.ones_like(rgb) * 0.3, torch.ones_like(depth)])
        res.append(gr)
    # Importance sampling is discontinuous (searchsorted): a 1e-6 difference in a coarse weight can move a handful of the 196 608 fine
    # samples to another bin, so the norms are compared loosely and the typical element tightly.
    for i, (a, b) in enumerate(zip(*res)):
        assert torch.isfinite(a).all() and torch.isfinite(b).all()
        assert ((a - b).norm() / b.norm()).item() < (1e-2 if i == 0 else 5e-3), i
    a, b = res[0][0].flatten(), res[1][0].flatten()
    nzi = b.abs() > 1e-3 * b.abs().max()
    rel = ((a[nzi] - b[nzi]).abs() / b[nzi].abs())
    assert rel.median().item() < 1e-4 and (rel > 1e-2).float().mean().item() < 1e-3


def test_render_baseline_geometry_vs_oracle():
    """BASELINE geometry (one image: 512^2 x 32-ch tri-planes, 64x64 patch rays from the reference's ray generator, 48+48 samples/ray) against the CPU
    oracle run on this box: forward outputs within 1e-3 max-rel, plane / ray gradients within 1e-3 and MLP parameter gradients within 3e-3 l2-rel (training default: 3xTF32 MLP)."""
    rm = _rm()
    rs = np.random.RandomState(11)
    B, P, res, N = 1, 512, 64, 48
    Rr = res * res
    planes = rs.standard_normal((B, 3, 32, P, P)).astype(np.float32)
    w1 = rs.standard_normal((64, 32)).astype(np.float32); b1 = (0.2 * rs.standard_normal(64)).astype(np.float32)
    w2 = rs.standard_normal((4, 64)).astype(np.float32); b2 = (0.2 * rs.standard_normal(4)).astype(np.float32)
    angles = torch.tensor([[0.4, 1.3, 0.0]]); look = torch.tensor([[0.5, 1.4, 0.1]])
    c2w = R.compute_cam2world_matrix(angles, torch.ones(B), look)
    ro, rd = R.sample_rays(c2w, torch.tensor([28.0]), (res, res), torch.full((B, 2), 0.5), torch.tensor([[0.2, 0.3]]))
    u1 = torch.from_numpy(rs.uniform(0, 1, (B, Rr, N)).astype(np.float32)); u2 = torch.from_numpy(rs.uniform(0, 1, (B, Rr, N)).astype(np.float32))
    tp = [torch.from_numpy(a).requires_grad_(True) for a in (planes, w1, b1, w2, b2)]
    ro_c, rd_c = ro.clone().requires_grad_(True), rd.clone().requires_grad_(True)
    old = R.DIFFERENTIABLE
    R.DIFFERENTIABLE = True
    try:
        o_rgb, o_depth, o_wsum, o_T = R.render(*tp, ro_c, rd_c, u1, u2, 0.75, 1.25, 0.5, N)
        g_rgb = torch.from_numpy(cases.cotangent(tuple(o_rgb.shape), 21)); g_dep = torch.from_numpy(cases.cotangent(tuple(o_depth.shape), 22))
        og = torch.autograd.grad([o_rgb, o_depth], tp + [ro_c, rd_c], [g_rgb, g_dep])
    finally:
        R.DIFFERENTIABLE = old
    args = [cu(a).requires_grad_(True) for a in (planes, w1, b1, w2, b2)] + [ro.cuda().requires_grad_(True), rd.cuda().requires_grad_(True)]
    rgb, depth, wsum, tfin = rm.render_rays(*args, num_steps=N, ray_start=0.75, ray_end=1.25, box_size=1.0, u_coarse=u1.cuda(), u_fine=u2.cuda(), mlp_mode=2)
    assert maxrel(rgb.detach().cpu().numpy(), o_rgb.detach().numpy()) < TOL
    assert maxrel(depth.detach().squeeze(-1).cpu().numpy(), o_depth.detach().numpy()) < TOL
    assert maxrel(wsum.detach().squeeze(-1).cpu().numpy(), o_wsum.detach().numpy()) < TOL
    assert maxrel(tfin.detach().cpu().numpy(), o_T.detach().numpy()) < TOL
    gs = torch.autograd.grad([rgb, depth], args, [g_rgb.cuda(), g_dep.cuda().unsqueeze(-1)])
    # plane / ray gradients 1e-3; MLP parameter gradients are sums over 393 216 samples with heavy cancellation (fp32 on both sides): 3e-3 as for the
    # network-level parameter gradients
    errs = {nm: l2rel(a.contiguous().cpu().numpy().reshape(-1), b.numpy().reshape(-1)) for nm, a, b in zip(['planes', 'w1', 'b1', 'w2', 'b2', 'ray_o', 'ray_d'], gs, og)}
    assert max(errs[k] for k in ('planes', 'ray_o', 'ray_d')) < TOL and max(errs[k] for k in ('w1', 'b1', 'w2', 'b2')) < 3e-3, errs


def _camera(B, seed=3):
    rs = np.random.RandomState(seed)
    dn = importlib.import_module('3dgp_b200.dnnlib')
    ru = importlib.import_module('3dgp_b200.training.rendering_utils')
    angles = torch.from_numpy(np.stack([rs.uniform(-1.2, 1.2, B), rs.uniform(0.9, 2.2, B), np.zeros(B)], 1).astype(np.float32)).cuda()
    look = torch.from_numpy(np.stack([rs.uniform(-3, 3, B), rs.uniform(0.2, 2.9, B), rs.uniform(0, 0.2, B)], 1).astype(np.float32)).cuda()
    fov = torch.from_numpy(rs.uniform(12, 40, B).astype(np.float32)).cuda()
    c2w = ru.compute_cam2world_matrix(dn.TensorGroup(angles=angles, radius=torch.ones(B, device='cuda'), look_at=look))
    return c2w, fov


@pytest.mark.parametrize('res,patch', [((16, 16), True), ((6, 10), False), ((9, 7), True)])
def test_in_kernel_ray_generation_matches_sample_rays(res, patch):
    """gp3d_generate_rays and the camera-fused forward (4 x 4 pixel tiles, partial tiles at the image border) against the module-level ray generator
    (tri_plane_renderer.py:487-527 restated in torch) feeding the explicit-ray entry point."""
    rm = _rm()
    tpr = importlib.import_module('3dgp_b200.training.tri_plane_renderer')
    B, P, N = 2, 32, 12
    h, w = res
    rs = np.random.RandomState(5)
    c2w, fov = _camera(B)
    ps = torch.full((B, 2), 0.5, device='cuda') if patch else None
    po = torch.tensor([[0.25, 0.375], [0.1, 0.3]], device='cuda') if patch else None
    pp = dict(scales=ps, offsets=po) if patch else None
    ro_t, rd_t = tpr.sample_rays(c2w, fov=fov, resolution=(w, h), patch_params=pp, device='cuda')     # the reference unpacks `w, h = resolution` (:497)
    ro_k, rd_k = rm.generate_rays(c2w, fov, (h, w), ps, po)
    assert maxrel(ro_k.cpu().numpy(), ro_t.cpu().numpy()) < 1e-6 and maxrel(rd_k.cpu().numpy(), rd_t.cpu().numpy()) < 2e-6
    planes = cu(rs.standard_normal((B, 3, 32, P, P)).astype(np.float32))
    w1 = cu(rs.standard_normal((64, 32)).astype(np.float32)); b1 = cu((0.2 * rs.standard_normal(64)).astype(np.float32))
    w2 = cu(rs.standard_normal((4, 64)).astype(np.float32)); b2 = cu((0.2 * rs.standard_normal(4)).astype(np.float32))
    u1 = cu(rs.uniform(0, 1, (B, h * w, N)).astype(np.float32)); u2 = cu(rs.uniform(0, 1, (B, h * w, N)).astype(np.float32))
    kw = dict(num_steps=N, ray_start=0.75, ray_end=1.25, box_size=1.0, u_coarse=u1, u_fine=u2)
    a = rm.render_camera(planes, w1, b1, w2, b2, c2w, fov, (h, w), ps, po, mlp_mode=2, **kw)
    b = rm.render_rays(planes, w1, b1, w2, b2, ro_t, rd_t, mlp_mode=0, **kw)            # first-generation fp32 SIMT kernel, explicit rays
    for x, y in zip(a, b):
        assert maxrel(x.cpu().numpy(), y.cpu().numpy()) < 1e-4


def test_camera_render_gradients_reach_the_camera():
    """d(rgb, depth) / d(cam2world, fov) of the camera-fused op == autograd through sample_rays + the explicit-ray op (the camera adaptor's path)."""
    rm = _rm()
    tpr = importlib.import_module('3dgp_b200.training.tri_plane_renderer')
    B, P, N, h, w = 2, 32, 8, 8, 8
    rs = np.random.RandomState(9)
    c2w, fov = _camera(B, seed=4)
    # smooth planes (4 x 4 noise, bilinearly magnified): d(output)/d(ray) over white-noise texels is a sum of large random terms whose cancellation
    # amplifies the last-bit differences between the two ray generators far beyond either path's own error
    coarse = torch.from_numpy(rs.standard_normal((B * 3, 32, 4, 4)).astype(np.float32))
    planes = torch.nn.functional.interpolate(coarse, size=(P, P), mode='bilinear', align_corners=True).reshape(B, 3, 32, P, P).cuda()
    w1 = cu(rs.standard_normal((64, 32)).astype(np.float32)); b1 = cu(np.zeros(64, np.float32))
    w2 = cu(rs.standard_normal((4, 64)).astype(np.float32)); b2 = cu(np.zeros(4, np.float32))
    u1 = cu(rs.uniform(0, 1, (B, h * w, N)).astype(np.float32)); u2 = cu(rs.uniform(0, 1, (B, h * w, N)).astype(np.float32))
    ps = torch.full((B, 2), 0.5, device='cuda'); po = torch.full((B, 2), 0.25, device='cuda')
    kw = dict(num_steps=N, ray_start=0.75, ray_end=1.25, box_size=1.0, u_coarse=u1, u_fine=u2, mlp_mode=2)
    g1 = cu(cases.cotangent((B, h * w, 3), 5)); g2 = cu(cases.cotangent((B, h * w, 1), 6))
    outs = []
    for fused in (True, False):
        c = c2w.detach().clone().requires_grad_(True); f = fov.detach().clone().requires_grad_(True); pl = planes.clone().requires_grad_(True)
        if fused:
            rgb, depth, _, _ = rm.render_camera(pl, w1, b1, w2, b2, c, f, (h, w), ps, po, **kw)
        else:
            ro, rd = tpr.sample_rays(c, fov=f, resolution=(h, w), patch_params=dict(scales=ps, offsets=po), device='cuda')
            rgb, depth, _, _ = rm.render_rays(pl, w1, b1, w2, b2, ro, rd, **kw)
        outs.append((rgb, depth) + torch.autograd.grad([rgb, depth], [c, f, pl], [g1, g2]))
    # outputs 1e-4; gradients at the suite's gradient bar (1e-3 l2-rel): the two ray generators differ in the last bit, and the inverse-CDF fine
    # samples over white-noise planes amplify that
    for i, (x, y) in enumerate(zip(*outs)):
        assert l2rel(x.detach().cpu().numpy(), y.detach().cpu().numpy()) < (1e-4 if i < 2 else 1e-3), i
    assert outs[0][2].abs().max() > 0 and outs[0][3].abs().max() > 0


def test_philox_stream_is_shared_by_all_kernel_generations():
    """Production mode (in-kernel Philox jitter, inverse-CDF uniforms and density noise): the third-generation forward must draw exactly the variates
    the first-generation kernel draws -- the backward kernels regenerate them from (seed, offset, ray, sample, pass)."""
    rm = _rm()
    rs = np.random.RandomState(2)
    B, P, N, Rr = 2, 32, 12, 50
    inp = cases.render_inputs('philox', dict(B=B, R=Rr, N=N, P=P, C=32, H=64))
    t = {k: cu(v) for k, v in inp.items()}
    kw = dict(num_steps=N, ray_start=0.75, ray_end=1.25, box_size=1.0, density_noise=0.7, seed=11, offset=3)
    a = rm.render_rays(t['planes'], t['w1'], t['b1'], t['w2'], t['b2'], t['ray_o'], t['ray_d'], mlp_mode=2, **kw)
    b = rm.render_rays(t['planes'], t['w1'], t['b1'], t['w2'], t['b2'], t['ray_o'], t['ray_d'], mlp_mode=0, **kw)
    for x, y in zip(a, b):
        assert maxrel(x.cpu().numpy(), y.cpu().numpy()) < 2e-4
