"""Op-level overlay of INTEGRATION.md (section B, first two rows), executed: the UNMODIFIED reference network code (src/training/{networks_epigraf,
networks_stylegan2, networks_discriminator, networks_depth_adaptor, layers}.py, imported in place from /root/reference) with THIS repo's op modules
(3dgp_b200/torch_utils/ops/{bias_act, upfirdn2d, conv2d_resample, conv2d_gradfix, fma}.py) bound under the names the reference imports
(`from src.torch_utils.ops import ...`, networks_stylegan2.py:21-24, layers.py:9-11, networks_discriminator.py:7) -- i.e. every call site of the
reference exercises the product ops' public surface: `conv2d_resample(x=, w=, f=, up=, down=, padding=, groups=, flip_weight=)`, `bias_act(x, b, act=,
gain=, clamp=)`, `activation_funcs[...]`, `setup_filter`, `upsample2d`, `fma`, `conv2d_gradfix.no_weight_gradients()`.

Results are held to the goldens the same reference code produced with its OWN ops (tests/golden/networks*.npz).  No GPU here: the product ops run on the
emulated C ABI (tests/abi_emulator.py); skipped where /root/reference is absent."""
import importlib
import json
import os

import numpy as np
import pytest
import torch

import abi_emulator as emu
from conftest import ROOT
from oracle import cases, ref_harness as rh
from util import l2rel, maxrel

pytestmark = pytest.mark.skipif(not rh.available(), reason='the unmodified reference is only present in the build container')
OPS = ('bias_act', 'upfirdn2d', 'conv2d_resample', 'conv2d_gradfix', 'fma')


@pytest.fixture()
def overlay(monkeypatch):
    tc = emu.install(monkeypatch)
    ns = rh.load()
    ours = {n: importlib.import_module('3dgp_b200.torch_utils.ops.' + n) for n in OPS}
    import src.training.networks_depth_adaptor as ref_da
    import src.training.networks_camera_adaptor as ref_ca
    bound = 0
    for mod in (ns.networks_epigraf, ns.networks_stylegan2, ns.networks_discriminator, ns.layers, ref_da, ref_ca):
        for n, o in ours.items():
            if hasattr(mod, n):
                monkeypatch.setattr(mod, n, o); bound += 1
    assert bound == 8          # networks_stylegan2.py:21-24 (4), layers.py:9-11 (3), networks_discriminator.py:7 (1)
    return ns, tc, ours


def _build(ns, variant):
    kw = cases.net_kwargs(variant)
    Gc, Dc, m = rh.make_cfg(**kw)
    G = rh.build_reference_G(Gc, m['img_resolution'], seed=0)
    D = rh.build_reference_D(Dc, m['patch_res'], use_depth=True, embedding_dim=m['embedding_dim'], seed=1, fp32=True)
    G.load_state_dict(cases.fill_state_dict({k: tuple(v.shape) for k, v in G.state_dict().items()}, G.state_dict(), seed=100))
    D.load_state_dict(cases.fill_state_dict({k: tuple(v.shape) for k, v in D.state_dict().items()}, D.state_dict(), seed=200))
    t = {k: torch.from_numpy(v) for k, v in cases.net_inputs(kw).items()}
    cam = ns.dnnlib.TensorGroup(angles=t['angles'], fov=t['fov'], radius=t['radius'], look_at=t['look_at'])
    return kw, G, D, t, cam, dict(scales=t['patch_scales'], offsets=t['patch_offsets'])


def test_reference_networks_on_our_ops_small(overlay, monkeypatch):
    ns, tc, ours = overlay
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'networks.npz'))
    kw, G, D, t, cam, pp = _build(ns, 'small')
    B, N, Rr = t['z'].shape[0], kw['num_ray_steps'], kw['patch_res'] ** 2
    noises = [torch.from_numpy(n) for n in cases.layer_noises(kw, B)]
    head = cases.depth_heads(B)
    monkeypatch.setattr(np.random, 'choice', lambda *a, **k: head.copy())
    s0 = dict(tc.stats)
    # (training forward and Gmain gradients of this overlay: test_reference_synthesis_network_on_our_renderer below)
    with torch.no_grad():
        ws = G.mapping(t['z'], t['c'])
    assert maxrel(ws.numpy(), gold['G/ws']) < 1e-5
    D.train()
    # eval: the grouped-conv (fused_modconv) form of modulated_conv2d, const noise
    G.eval()
    ue = cases.eval_variates(kw, B); Re = kw['img_resolution'] ** 2
    with torch.no_grad(), rh.injected_rng(rand_like=[torch.from_numpy(ue['u_coarse']).reshape(B, Re, N, 1)], rand=[torch.from_numpy(ue['u_fine']).reshape(B * Re, N)]):
        oe = G.synthesis(ws.detach(), cam, noise_mode='const', render_opts=dict(concat_depth=True, return_depth=True))
    assert maxrel(oe.img.numpy(), gold['G/eval/img']) < 1e-4
    assert tc.stats['aten'] > s0['aten'], 'the convolutions of the reference modules went through the product conv2d_gradfix'
    # D forward + the R1 double backward under the PRODUCT's no_weight_gradients (loss.py:238-253 uses conv2d_gradfix.no_weight_gradients)
    img = torch.from_numpy(gold['G/train/img']).requires_grad_(True)
    logits, feats = D(img, t['c'], patch_params=pp, camera_angles=t['angles'], predict_feat=True)
    assert maxrel(logits.detach().numpy(), gold['D/logits']) < 1e-4 and maxrel(feats.detach().numpy(), gold['D/feats']) < 1e-4
    with ours['conv2d_gradfix'].no_weight_gradients():
        r1 = torch.autograd.grad([logits.sum()], [img], create_graph=True)[0]
    assert l2rel(r1.detach().numpy(), gold['D/r1_grads']) < 1e-4
    loss = torch.nn.functional.softplus(-logits).mean() + r1.square().sum([1, 2, 3]).mean() * 0.5
    names = cases.probe_params('D'); pars = dict(D.named_parameters())
    for n, gr in zip(names, torch.autograd.grad(loss, [pars[n] for n in names])):
        assert l2rel(gr.numpy(), gold['D/grad/' + n]) < 5e-4, n


def test_reference_decoder_and_discriminator_on_our_tensor_core_route_wide(overlay):
    """Same overlay at tensor-core-eligible widths: the reference modules' convolutions now take the product's tensor-core PRIMITIVES (conv2d_gradfix ->
    ops/tc.py tap lists; the fused layer nodes belong to the module-level overlay and are not involved here)."""
    ns, tc, ours = overlay
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'networks_wide.npz'))
    kw, G, D, t, cam, pp = _build(ns, 'wide')
    B = t['z'].shape[0]
    noises = [torch.from_numpy(n) for n in cases.layer_noises(kw, B)]
    ws = torch.from_numpy(gold['G/ws'])
    G.train()
    s0 = dict(tc.stats)
    with torch.no_grad(), rh.injected_rng(randn=[n.clone() for n in noises]):
        planes = G.synthesis.tri_plane_decoder(ws, noise_mode='random', fused_modconv=False)
    assert tc.stats['tc'] - s0['tc'] >= 14 and tc.stats['fused'] == s0['fused']
    assert maxrel(planes.flatten()[::31].numpy(), gold['G/train/planes_probe']) < 1e-4
    stats = np.array([planes.double().sum().item(), planes.double().abs().sum().item()])
    assert np.abs(stats - gold['G/train/planes_stats']).max() < 1e-4 * np.abs(gold['G/train/planes_stats']).max()
    D.train()
    s0 = dict(tc.stats)
    img = torch.from_numpy(gold['G/train/img']).requires_grad_(True)
    logits, feats = D(img, t['c'], patch_params=pp, camera_angles=t['angles'], predict_feat=True)
    assert tc.stats['tc'] - s0['tc'] >= 10
    assert maxrel(logits.detach().numpy(), gold['D/logits']) < 1e-4 and maxrel(feats.detach().numpy(), gold['D/feats']) < 1e-4
    with ours['conv2d_gradfix'].no_weight_gradients():
        r1 = torch.autograd.grad([logits.sum()], [img], create_graph=True)[0]
    assert l2rel(r1.detach().numpy(), gold['D/r1_grads']) < 5e-4


def test_reference_synthesis_network_on_our_renderer(overlay, monkeypatch):
    """`src/training/tri_plane_renderer.py` row of INTEGRATION.md: the unmodified reference SynthesisNetwork.forward (networks_epigraf.py:210-262) with THIS repo's
    ImportanceRenderer in place of its own -- called positionally `(planes, decoder, ray_o, ray_d, rendering_options)` in training and, for frames above
    `max_batch_res`, through `run_batchwise(fn=self.renderer, data=dict(ray_origins=, ray_directions=), dim=1, planes=, decoder=, rendering_options=)`
    (:232-240) in ray chunks; the decoder it hands over is the reference's own TriPlaneMLP.  Held to the reference goldens (image, depth, Gmain gradients)."""
    ns, tc, ours = overlay
    tpr = importlib.import_module('3dgp_b200.training.tri_plane_renderer')
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'networks.npz'))
    kw, G, D, t, cam, pp = _build(ns, 'small')
    B, N = t['z'].shape[0], kw['num_ray_steps']

    class Injecting(tpr.ImportanceRenderer):
        """Supplies the goldens' sampling variates (the reference draws them with torch.rand inside its renderer; the product takes them as options or from its
        counter-based stream) -- sliced along the ray axis when the caller renders in chunks."""
        def arm(self, u_coarse, u_fine):
            self.u, self.cursor, self.calls = (u_coarse, u_fine), 0, 0

        def forward(self, planes, decoder, ray_origins, ray_directions, rendering_options):
            R = ray_origins.shape[1]
            ro = dict(rendering_options, mlp_mode=0, u_coarse=self.u[0][:, self.cursor:self.cursor + R].contiguous(), u_fine=self.u[1][:, self.cursor:self.cursor + R].contiguous())
            self.cursor += R; self.calls += 1
            return super().forward(planes, decoder, ray_origins, ray_directions, ns.dnnlib.EasyDict(**ro))
    G.synthesis.renderer = Injecting('classical')
    noises = [torch.from_numpy(n) for n in cases.layer_noises(kw, B)]
    head = cases.depth_heads(B)
    monkeypatch.setattr(np.random, 'choice', lambda *a, **k: head.copy())
    G.train(); G.synthesis.nerf_noise_std = 0.0
    G.synthesis.renderer.arm(t['u_coarse'], t['u_fine'])
    with rh.injected_rng(randn=[n.clone() for n in noises]):
        ws = G.mapping(t['z'], t['c'])
        o = G.synthesis(ws, cam, patch_params=pp, render_opts=dict(concat_depth=True, return_depth=True))
    assert G.synthesis.renderer.calls == 1
    assert maxrel(o.img.detach().numpy(), gold['G/train/img']) < 1e-4 and maxrel(o.depth.detach().numpy(), gold['G/train/depth']) < 1e-4
    D.train()
    logits, _ = D(o.img, t['c'], patch_params=pp, camera_angles=t['angles'])
    loss = torch.nn.functional.softplus(-logits).mean()
    assert abs(loss.item() - float(gold['G/loss'][0])) < 1e-4
    names = cases.probe_params('G'); pars = dict(G.named_parameters())
    for n, gr in zip(names, torch.autograd.grad(loss, [pars[n] for n in names])):
        assert l2rel(gr.numpy(), gold['G/grad/' + n]) < 5e-4, n
    # eval, full frame rendered in ray chunks through run_batchwise's keyword call
    G.eval()
    ue = cases.eval_variates(kw, B)
    G.synthesis.renderer.arm(torch.from_numpy(ue['u_coarse']), torch.from_numpy(ue['u_fine']))
    with torch.no_grad():
        oe = G.synthesis(ws.detach(), cam, noise_mode='const', render_opts=dict(concat_depth=True, return_depth=True, max_batch_res=8))
    R, chunk = kw['img_resolution'] ** 2, N * 8 * 8                 # run_batchwise cuts the ray axis into chunks of num_ray_steps * max_batch_res^2 (:236)
    assert G.synthesis.renderer.calls == -(-R // chunk) > 1 and G.synthesis.renderer.cursor == R
    assert maxrel(oe.img.numpy(), gold['G/eval/img']) < 1e-4 and maxrel(oe.depth.numpy(), gold['G/eval/depth']) < 1e-4


def test_reference_module_summary_walks_our_networks(monkeypatch, capsys):
    """Start-up step of the reference loop (training_loop.py:140-160): `misc.print_module_summary` hooks every sub-module of G and D, runs one forward with the
    loop's own arguments (z, c, camera_params from `sample_camera_params`; the patch-sized image, zero patch parameters and camera angles for D) and tabulates
    parameters / buffers / output shapes -- on THIS repo's modules."""
    emu.install(monkeypatch)
    ns = rh.load()
    from src.torch_utils import misc
    from src.training.rendering_utils import sample_camera_params
    cfgm = importlib.import_module('3dgp_b200.config')
    kw = {k: v for k, v in cases.net_kwargs('small').items() if k != 'learn_camera_dist'}
    kw.update(img_resolution=16, patch_res=8, tri_res=16, num_ray_steps=4)          # the walk is about module surfaces, not sizes
    cfg = cfgm.make_config(**kw, kd_weight=1.0, batch_size=4)
    G, D = cfgm.build_networks(cfg, 'cpu', fp32_D=True)
    tb = 2
    with torch.no_grad():
        G.eval(); D.eval()
        z, c, angles = torch.randn([tb, G.z_dim]), torch.zeros([tb, G.c_dim]), torch.zeros([tb, 3])
        cam = sample_camera_params(ns.dnnlib.EasyDict.init_recursively(json.loads(json.dumps(G.cfg.camera))), tb, 'cpu', angles)
        img = misc.print_module_summary(G, [z[[0]], c[[0]]], module_kwargs={'camera_params': cam[[0]]})
        assert torch.is_tensor(img) and tuple(img.shape) == (1, 3, kw['img_resolution'], kw['img_resolution'])
        img = img.repeat(tb, 1, 1, 1)[:, :, :cfg.training.patch.resolution, :cfg.training.patch.resolution][:, [0]].repeat(1, 4, 1, 1)
        logits = misc.print_module_summary(D, [img, c], module_kwargs={'patch_params': {'scales': torch.zeros(tb, 2), 'offsets': torch.zeros(tb, 2)}, 'camera_angles': torch.zeros(tb, 3)})
    text = capsys.readouterr().out
    n_g, n_d = sum(p.numel() for p in G.parameters()), sum(p.numel() for p in D.parameters())
    assert str(n_d) in text and 'synthesis.tri_plane_decoder.b16:0' in text and 'synthesis.depth_adaptor.head' in text and 'b4.mbstd' in text      # every visited sub-module is listed; D's total is its parameter count
    assert n_g > 0 and tuple(logits[0].shape) == (tb,)                # D returns (logits, features)


def test_epigraf_model_configuration_matches_the_reference_run_live(monkeypatch):
    """The second 3-D model family of the reference, configs/model/epigraf.yaml (no depth adaptor, no depth channel into D, `discriminator.fmaps: 0.5`): the
    UNMODIFIED reference networks run live on CPU next to this repo's modules (emulated ABI) with the same weights, layer noise and sampling variates --
    training forward, Gmain loss through D, D logits."""
    emu.install(monkeypatch)
    ns = rh.load()
    cfgm = importlib.import_module('3dgp_b200.config')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    kw = {k: v for k, v in cases.net_kwargs('small').items() if k != 'learn_camera_dist'}
    kw.update(use_depth=False, d_fmaps=0.5)
    Gc, Dc, m = rh.make_cfg(**kw)
    Gr = rh.build_reference_G(Gc, m['img_resolution'], seed=0)
    Dr = rh.build_reference_D(Dc, m['patch_res'], use_depth=False, embedding_dim=0, seed=1, fp32=True)
    cfg = cfgm.make_config(**kw, kd_weight=0.0)
    G, D = cfgm.build_networks(cfg, 'cpu', fp32_D=True)
    assert G.synthesis.depth_adaptor is None and Gr.synthesis.depth_adaptor is None
    assert {k: tuple(v.shape) for k, v in G.state_dict().items()} == {k: tuple(v.shape) for k, v in Gr.state_dict().items()}
    assert {k: tuple(v.shape) for k, v in D.state_dict().items()} == {k: tuple(v.shape) for k, v in Dr.state_dict().items()}
    sdG = cases.fill_state_dict({k: tuple(v.shape) for k, v in Gr.state_dict().items()}, Gr.state_dict(), seed=100)
    sdD = cases.fill_state_dict({k: tuple(v.shape) for k, v in Dr.state_dict().items()}, Dr.state_dict(), seed=200)
    for net, sd in ((Gr, sdG), (G, sdG), (Dr, sdD), (D, sdD)):
        net.load_state_dict(sd); net.train()
    t = {k: torch.from_numpy(v) for k, v in cases.net_inputs(kw).items()}
    B, N, Rr = t['z'].shape[0], kw['num_ray_steps'], kw['patch_res'] ** 2
    noises = [torch.from_numpy(n) for n in cases.layer_noises(kw, B)]
    pp = dict(scales=t['patch_scales'], offsets=t['patch_offsets'])
    Gr.synthesis.nerf_noise_std = G.synthesis.nerf_noise_std = 0.0
    with torch.no_grad():
        cam_r = ns.dnnlib.TensorGroup(angles=t['angles'], fov=t['fov'], radius=t['radius'], look_at=t['look_at'])
        with rh.injected_rng(randn=[n.clone() for n in noises], rand_like=[t['u_coarse'].reshape(B, Rr, N, 1)], rand=[t['u_fine'].reshape(B * Rr, N)]):
            img_r = Gr.synthesis(Gr.mapping(t['z'], t['c']), cam_r, patch_params=pp)
        logits_r, _ = Dr(img_r, t['c'], patch_params=pp, camera_angles=t['angles'])
        cam = dn.TensorGroup(angles=t['angles'], fov=t['fov'], radius=t['radius'], look_at=t['look_at'])
        img = G.synthesis(G.mapping(t['z'], t['c']), cam, patch_params=pp, noise_mode='random', layer_noises=noises,
                          render_opts=dict(u_coarse=t['u_coarse'], u_fine=t['u_fine'], mlp_mode=0))
        logits, _ = D(img, t['c'], patch_params=pp, camera_angles=t['angles'])
    assert torch.is_tensor(img) and tuple(img.shape) == tuple(img_r.shape) == (B, 3, kw['patch_res'], kw['patch_res'])
    assert maxrel(img.numpy(), img_r.numpy()) < 1e-4 and maxrel(logits.numpy(), logits_r.numpy()) < 1e-4


def test_reference_augment_pipe_on_our_ops(monkeypatch):
    """src/training/augment.py (ADA, off in the 3dgp configuration but part of the reference loop) imports `upfirdn2d` and `conv2d_gradfix` from the overlaid
    package (:19-21) and calls forms nothing else on the path uses: `upsample2d` / `downsample2d` with a 12-tap separable wavelet filter, negative padding
    and `flip_filter=True` (:294, :305), and `conv2d` with `groups = batch * channels` (:413-414).  All fifteen augmentations on, same RNG state: output and
    input gradient with this repo's op modules bound equal the reference's own."""
    ns = rh.load()
    import src.training.augment as aug
    pipe = aug.AugmentPipe(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1, brightness=1, contrast=1, lumaflip=1, hue=1, saturation=1,
                           imgfilter=1, noise=1, cutout=1).train()
    x = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(0))

    def run():
        torch.manual_seed(1); np.random.seed(1)
        xi = x.clone().requires_grad_(True)
        y = pipe(xi, num_color_channels=3)
        g, = torch.autograd.grad(y.square().sum(), xi)
        return y.detach(), g
    y_ref, g_ref = run()                                  # the reference's own ops (their CPU implementations)
    emu.install(monkeypatch)
    monkeypatch.setattr(aug, 'upfirdn2d', importlib.import_module('3dgp_b200.torch_utils.ops.upfirdn2d'))
    monkeypatch.setattr(aug, 'conv2d_gradfix', importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix'))
    y, g = run()
    assert maxrel(y.numpy(), y_ref.numpy()) < 1e-5 and maxrel(g.numpy(), g_ref.numpy()) < 1e-5


def test_compute_densities_matches_the_reference_run_live(monkeypatch):
    """`SynthesisNetwork.compute_densities` (networks_epigraf.py:196-208, the call scripts/extract_geometry.py:30 makes): this repo's method next to the
    reference's on the same weights and query points (inside and outside the scene cube), evaluated in several chunks."""
    emu.install(monkeypatch)
    ns = rh.load()
    cfgm = importlib.import_module('3dgp_b200.config')
    kw = {k: v for k, v in cases.net_kwargs('small').items() if k != 'learn_camera_dist'}
    Gc, _, m = rh.make_cfg(**kw)
    Gr = rh.build_reference_G(Gc, m['img_resolution'], seed=0).eval()
    G, _ = cfgm.build_networks(cfgm.make_config(**kw), 'cpu', fp32_D=True)
    sd = cases.fill_state_dict({k: tuple(v.shape) for k, v in Gr.state_dict().items()}, Gr.state_dict(), seed=100)
    Gr.load_state_dict(sd); G.load_state_dict(sd); G.eval()
    t = {k: torch.from_numpy(v) for k, v in cases.net_inputs(kw).items()}
    g = torch.Generator().manual_seed(4)
    coords = (torch.rand(t['z'].shape[0], 700, 3, generator=g) * 2 - 1) * 0.6            # cube half-extent is 0.5: some points fall outside
    with torch.no_grad():
        ws = Gr.mapping(t['z'], t['c'])
        want = Gr.synthesis.compute_densities(ws, coords, max_batch_res=6, noise_mode='const')
        got = G.synthesis.compute_densities(ws, coords, max_batch_res=6, noise_mode='const')
    assert tuple(got.shape) == tuple(want.shape) == (t['z'].shape[0], 700, 1)
    assert maxrel(got.numpy(), want.numpy()) < 1e-4


def test_stylegan2_2d_generator_matches_the_reference_run_live(monkeypatch):
    """`model=stylegan2` (the 2-D generator of networks_stylegan2.py:281-375, kept so the overlaid file serves every model): this repo's classes next to the
    reference's, same weights and layer noise, fp32 blocks; training forward with patch extraction and the eval (grouped-conv) form."""
    emu.install(monkeypatch)
    ns = rh.load()
    sg = importlib.import_module('3dgp_b200.training.networks_stylegan2')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    g_cfg = dict(cbase=512, cmax=32, fmaps=1.0, w_dim=32, z_dim=32, c_dim=4, map_depth=2, architecture='skip',
                 patch=dict(enabled=True, resolution=16), camera_cond=False)
    kwargs = dict(img_resolution=32, img_channels=3, mapping_kwargs=dict(camera_cond=False, camera_cond_drop_p=0.0, mean_camera_params=None),
                  fused_modconv_default='inference_only', num_fp16_res=0, conv_clamp=None)
    torch.manual_seed(0)
    Gr = ns.networks_stylegan2.Generator(cfg=ns.dnnlib.EasyDict.init_recursively(g_cfg), **kwargs)
    G = sg.Generator(cfg=dn.EasyDict.init_recursively(g_cfg), **kwargs)
    assert {k: tuple(v.shape) for k, v in G.state_dict().items()} == {k: tuple(v.shape) for k, v in Gr.state_dict().items()} and G.num_ws == Gr.num_ws
    sd = cases.fill_state_dict({k: tuple(v.shape) for k, v in Gr.state_dict().items()}, Gr.state_dict(), seed=300)
    Gr.load_state_dict(sd); G.load_state_dict(sd)
    gen = torch.Generator().manual_seed(2)
    z, c = torch.randn(4, 32, generator=gen), torch.eye(4)
    pp = dict(scales=torch.full([4, 2], 0.5), offsets=torch.rand(4, 2, generator=gen) * 0.5)
    noises = [torch.randn(4, 1, r, r, generator=gen) for r in (4, 8, 8, 16, 16, 32, 32)]
    Gr.train(); G.train()
    with torch.no_grad():
        with rh.injected_rng(randn=[n.clone() for n in noises]):
            want = Gr(z, c, patch_params=pp, render_opts=dict(return_depth=True))
        got = G(z, c, patch_params=pp, render_opts=dict(return_depth=True), noise_mode='random', layer_noises=noises)
        assert tuple(got.img.shape) == tuple(want.img.shape) == (4, 3, 16, 16) and maxrel(got.img.numpy(), want.img.numpy()) < 1e-4 and float(got.depth.abs().max()) == 0.0
        Gr.eval(); G.eval()
        assert maxrel(G(z, c, noise_mode='const').numpy(), Gr(z, c, noise_mode='const').numpy()) < 1e-4
