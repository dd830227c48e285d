"""Host wrapper of the fused ray-march (3dgp_b200/torch_utils/ops/raymarch.py) on a machine without a GPU: the four C-ABI entry points are served by the
oracle's renderer through the pointers / strides / option structure the wrapper hands over (tests/abi_emulator.py), and the results are compared with
the goldens the UNMODIFIED reference produced (tests/golden/render.npz: tri_plane_renderer.py:126-170 and its autograd).  What this pins without a GPU:
the channel-minor plane layout and the strides that describe it, option codes, output shapes, the routing of the seven gradients, the in-kernel
camera path (cam2world / fov / patch transform -> rays) and its chain to d(cam2world), d(fov)."""
import contextlib
import importlib
import os

import numpy as np
import pytest
import torch

import abi_emulator as emu
from conftest import ROOT
from oracle import cases, restated as R
from util import maxrel, l2rel

rm = importlib.import_module('3dgp_b200.torch_utils.ops.raymarch')
_lib = importlib.import_module('3dgp_b200._lib')


@pytest.fixture(autouse=True)
def emulated_abi(monkeypatch):
    fake = emu.FakeLib()
    monkeypatch.setattr(_lib, 'lib', lambda: fake)
    monkeypatch.setattr(_lib, 'stream_ptr', lambda: None)
    monkeypatch.setattr(_lib, 'require_cuda', lambda t, name='tensor': None)
    monkeypatch.setattr(torch.cuda, 'device', lambda _d: contextlib.nullcontext())
    monkeypatch.setattr(R, 'DIFFERENTIABLE', True)


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'render.npz'))


def _run(kw, inp, layout, requires_grad=False):
    t = {k: torch.from_numpy(v) for k, v in inp.items()}
    planes = t['planes']
    if layout == 'channel_minor':       # what the tri-plane decoder emits: [B, P, P, 3C] storage behind the [B, 3, C, P, P] view
        planes = rm.planes_channel_minor(planes)
        assert planes.stride(2) == 1
    args = [planes] + [t[k] for k in ('w1', 'b1', 'w2', 'b2', 'ray_o', 'ray_d')]
    if requires_grad:
        args = [a.clone().requires_grad_(True) if i else a.detach().requires_grad_(True) for i, a in enumerate(args)]
    out = rm.render_rays(*args, num_steps=kw['N'], ray_start=kw['ray_start'], ray_end=kw['ray_end'], box_size=2 * kw['box_half'],
                         u_coarse=t['u_coarse'], u_fine=t['u_fine'], sn_coarse=t.get('sn_coarse'), sn_fine=t.get('sn_fine'),
                         density_noise=kw.get('noise_std', 0.0), use_inf_depth=kw.get('use_inf_depth', True), last_back=kw.get('last_back', False),
                         white_back_end_idx=kw.get('white_back_end_idx', 0), clamp_mode=kw.get('clamp_mode', 'softplus'), mlp_mode=0)
    return out, args


@pytest.mark.parametrize('layout', ['nchw', 'channel_minor'])
@pytest.mark.parametrize('name,kw', cases.render_cases(), ids=[c[0] for c in cases.render_cases()])
def test_render_wrapper_forward_and_gradients_vs_reference_golden(gold, name, kw, layout):
    inp = cases.render_inputs(name, kw)
    (rgb, depth, wsum, tfin), args = _run(kw, inp, layout, requires_grad=True)
    assert tuple(depth.shape) == (kw['B'], kw['R'], 1) and tuple(wsum.shape) == (kw['B'], kw['R'], 1) and tuple(tfin.shape) == (kw['B'], kw['R'])
    assert maxrel(rgb.detach().numpy(), gold[name + '/rgb']) < 2e-5 and maxrel(depth.detach().squeeze(-1).numpy(), gold[name + '/depth']) < 2e-5
    assert maxrel(wsum.squeeze(-1).numpy(), gold[name + '/wsum']) < 2e-5 and maxrel(tfin.numpy(), gold[name + '/tfinal']) < 2e-5
    assert not wsum.requires_grad and not tfin.requires_grad
    g_rgb = torch.from_numpy(cases.cotangent(rgb.shape, 21)); g_dep = torch.from_numpy(cases.cotangent(depth.shape, 22))
    grads = torch.autograd.grad([rgb, depth], args, [g_rgb, g_dep])
    for nm, gr in zip(['g_planes', 'g_w1', 'g_b1', 'g_w2', 'g_b2', 'g_ray_o', 'g_ray_d'], grads):
        if nm == 'g_planes' and (name + '/g_planes') not in gold.files:
            assert l2rel(gr.contiguous().flatten()[::97].numpy(), gold[name + '/g_planes_probe']) < 1e-4
            continue
        assert gr.shape == gold[name + '/' + nm].shape and l2rel(gr.contiguous().numpy(), gold[name + '/' + nm]) < 1e-4, nm


def test_camera_path_equals_explicit_rays_and_chains_to_the_camera():
    name, kw = cases.render_cases()[0]
    kw = dict(kw, R=36)                                   # 6 x 6 ray grid
    inp = cases.render_inputs(name, kw)
    t = {k: torch.from_numpy(v) for k, v in inp.items()}
    B = kw['B']
    rs = np.random.RandomState(3)
    ang = torch.tensor(rs.uniform([-1.0, 1.0, 0.0], [1.0, 2.0, 0.0], size=(B, 3)), dtype=torch.float32)
    c2w = R.compute_cam2world_matrix(ang, torch.ones(B), torch.tensor(rs.uniform(-0.1, 0.1, size=(B, 3)), dtype=torch.float32) * 0 + torch.tensor([[0.5, 1.5, 0.1]] * B))
    fov = torch.tensor(rs.uniform(15, 40, size=B), dtype=torch.float32)
    ps = torch.tensor(rs.uniform(0.4, 0.9, size=(B, 2)), dtype=torch.float32); po = torch.tensor(rs.uniform(0.0, 0.1, size=(B, 2)), dtype=torch.float32)
    common = dict(num_steps=kw['N'], ray_start=kw['ray_start'], ray_end=kw['ray_end'], box_size=2 * kw['box_half'], u_coarse=t['u_coarse'], u_fine=t['u_fine'])
    ro, rd = rm.generate_rays(c2w, fov, (6, 6), ps, po)
    ro_ref, rd_ref = R.sample_rays(c2w, fov, (6, 6), ps, po)
    assert maxrel(ro.numpy(), ro_ref.numpy()) < 1e-6 and maxrel(rd.numpy(), rd_ref.numpy()) < 1e-6
    w = [t[k] for k in ('w1', 'b1', 'w2', 'b2')]
    c2g, fvg = c2w.clone().requires_grad_(True), fov.clone().requires_grad_(True)
    a = rm.render_camera(t['planes'], *w, c2g, fvg, (6, 6), ps, po, **common, mlp_mode=0)
    b = rm.render_rays(t['planes'], *w, ro, rd, **common)
    assert maxrel(a[0].detach().numpy(), b[0].numpy()) < 1e-6 and maxrel(a[1].detach().numpy(), b[1].numpy()) < 1e-6
    # d(cam2world), d(fov): against autograd through the oracle's ray generator + renderer
    g_c2w, g_fov = torch.autograd.grad(a[0].square().sum() + a[1].sum(), [c2g, fvg])
    c2r, fvr = c2w.clone().requires_grad_(True), fov.clone().requires_grad_(True)
    ro_r, rd_r = R.sample_rays(c2r, fvr, (6, 6), ps, po)
    out = R.render(t['planes'], *w, ro_r, rd_r, t['u_coarse'], t['u_fine'], kw['ray_start'], kw['ray_end'], kw['box_half'], kw['N'])
    r_c2w, r_fov = torch.autograd.grad(out[0].square().sum() + out[1].sum(), [c2r, fvr])
    assert l2rel(g_c2w[:, :3].numpy(), r_c2w[:, :3].numpy()) < 1e-4 and l2rel(g_fov.numpy(), r_fov.numpy()) < 1e-4


def test_wrapper_rejects_malformed_inputs():
    name, kw = cases.render_cases()[0]
    inp = cases.render_inputs(name, kw)
    t = {k: torch.from_numpy(v) for k, v in inp.items()}
    common = dict(num_steps=kw['N'], ray_start=0.75, ray_end=1.25, box_size=1.0)
    with pytest.raises(RuntimeError, match='B\\*R\\*N'):
        rm.render_rays(t['planes'], t['w1'], t['b1'], t['w2'], t['b2'], t['ray_o'], t['ray_d'], u_coarse=t['u_coarse'][:, :, :-1], u_fine=t['u_fine'], **common)
    with pytest.raises(RuntimeError, match='2 layers'):
        rm.render_rays(t['planes'], t['w1'][:, :-1], t['b1'], t['w2'], t['b2'], t['ray_o'], t['ray_d'], u_coarse=t['u_coarse'], u_fine=t['u_fine'], **common)
    with pytest.raises(RuntimeError, match='float32 or float16'):
        rm.render_rays(t['planes'].double(), t['w1'], t['b1'], t['w2'], t['b2'], t['ray_o'], t['ray_d'], u_coarse=t['u_coarse'], u_fine=t['u_fine'], **common)


def test_launcher_geometry_helpers_equal_the_reference():
    """`get_ray_limits_box` / `validate_image_plane` (kept in the overlaid tri_plane_renderer.py because src/train.py:29,211 imports them): bit-identical
    to the reference's slab test on random, missing and axis-parallel rays, same verdicts on camera configurations either side of the limit."""
    import importlib
    import pytest
    import torch
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip('the unmodified reference is only present in the build container')
    ours = importlib.import_module('3dgp_b200.training.tri_plane_renderer')
    ref = rh.load().tri_plane_renderer
    g = torch.Generator().manual_seed(0)
    o = torch.randn(3, 50, 3, generator=g) * 1.5
    d = torch.nn.functional.normalize(torch.randn(3, 50, 3, generator=g), dim=-1)
    d[0, 0] = torch.tensor([1.0, 0.0, 0.0]); d[0, 1] = torch.tensor([0.0, 0.0, -1.0]); o[0, 1] = torch.tensor([0.2, 0.3, 2.0])
    for box in (1.0, 2.0, 4.0):
        a, b = ref.get_ray_limits_box(o, d, box), ours.get_ray_limits_box(o, d, box)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and 0 < int((a[0] >= 0).sum()) < a[0].numel()
    verdicts = [(fov, r, s, ours.validate_image_plane(fov, r, s, step=5e-2)) for fov, r, s in ((45.0, 1.0, 0.5), (10.0, 1.0, 0.5), (120.0, 3.0, 0.5), (30.0, 2.0, 0.5))]
    assert [v[3] for v in verdicts] == [True, True, False, False]
    assert all(v[3] == ref.validate_image_plane(v[0], v[1], v[2], step=5e-2) for v in verdicts)


def test_camera_prior_sampling_equals_the_reference_draw_for_draw():
    """`sample_camera_params` of the stand-alone rendering_utils.py vs the reference's (rendering_utils.py:72-156) from the same torch / numpy RNG state, for
    every distribution the shipped camera configs use (configs/camera/{base, uniform, ...}.yaml): uniform, normal, truncnorm, spherical_uniform angles;
    constant / uniform / truncnorm scalars; given origin angles."""
    import copy
    import importlib
    import numpy as np
    import pytest
    import torch
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip('the unmodified reference is only present in the build container')
    ns = rh.load()
    ref = ns.rendering_utils
    ours = importlib.import_module('3dgp_b200.training.rendering_utils')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    base = rh.make_cfg()[0]['camera']

    def variant(**edits):
        c = copy.deepcopy(base)
        for path, v in edits.items():
            node = c
            keys = path.split('__')
            for k in keys[:-1]:
                node = node[k]
            node[keys[-1]] = v
        return c
    cases_ = {'uniform': variant(), 'normal': variant(origin__angles__dist='normal'), 'truncnorm': variant(origin__angles__dist='truncnorm'),
              'spherical': variant(origin__angles__dist='spherical_uniform'), 'radius_uniform': variant(origin__radius=dict(dist='uniform', min=0.8, max=1.2)),
              'fov_truncnorm': variant(fov=dict(dist='truncnorm', mean=20.0, std=3.0, min=10.0, max=45.0))}
    for name, c in cases_.items():
        for given in (None, torch.full([7, 3], 0.25)):
            torch.manual_seed(3); np.random.seed(3)
            a = ref.sample_camera_params(ns.dnnlib.EasyDict.init_recursively(copy.deepcopy(c)), 7, 'cpu', given)
            torch.manual_seed(3); np.random.seed(3)
            b = ours.sample_camera_params(dn.EasyDict.init_recursively(copy.deepcopy(c)), 7, 'cpu', given)
            assert all(torch.equal(a[k], b[k]) for k in ('angles', 'fov', 'radius', 'look_at')), name
    ang = base['origin']['angles']
    assert ref.get_mean_angles_values(ns.dnnlib.EasyDict.init_recursively(ang)) == ours.get_mean_angles_values(dn.EasyDict.init_recursively(ang))
    assert ref.get_mean_sampling_value(ns.dnnlib.EasyDict.init_recursively(base['fov'])) == ours.get_mean_sampling_value(dn.EasyDict.init_recursively(base['fov']))


def test_patch_sampling_and_schedules_equal_the_reference_draw_for_draw():
    """Stand-alone training_utils.py vs the reference's (training_utils.py:8-214) from the same RNG state: patch parameters for the three shipped
    distributions (configs/training/patch_{beta, uniform, discrete_uniform}.yaml), patch extraction, class sampling, the linear schedule."""
    import importlib
    import numpy as np
    import pytest
    import torch
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip('the unmodified reference is only present in the build container')
    ns = rh.load()
    ref = ns.training_utils
    ours = importlib.import_module('3dgp_b200.training.training_utils')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    common = dict(min_scale=0.25, max_scale=1.0, mbstd_group_size=4, alpha=1.0, beta=0.5, discrete_support=[0.125, 0.25, 0.5, 1.0])
    for dist in ('beta', 'uniform', 'discrete_uniform'):
        pc = dict(common, distribution=dist)
        torch.manual_seed(5); np.random.seed(5)
        a = ref.sample_patch_params(8, ns.dnnlib.EasyDict(**pc))
        torch.manual_seed(5); np.random.seed(5)
        b = ours.sample_patch_params(8, dn.EasyDict(**pc))
        assert torch.equal(a['scales'], b['scales']) and torch.equal(a['offsets'], b['offsets']) and a['scales'].dtype == b['scales'].dtype, dist
    x = torch.randn(8, 4, 32, 32, generator=torch.Generator().manual_seed(0))
    assert torch.equal(ref.extract_patches(x, a, resolution=16), ours.extract_patches(x, b, resolution=16))
    torch.manual_seed(6); ca = ref.sample_random_c(9, 5, 'cpu')
    torch.manual_seed(6); cb = ours.sample_random_c(9, 5, 'cpu')
    assert torch.equal(ca, cb)
    for step in (-1, 0, 10, 250, 500, 501, 10_000):
        assert ref.linear_schedule(step, 1.0, 0.25, 500) == ours.linear_schedule(step, 1.0, 0.25, 500)
        assert ref.linear_schedule(step, 0.0, 1.0, 300, start_step=100) == ours.linear_schedule(step, 0.0, 1.0, 300, start_step=100)
