"""Whole networks on a machine WITHOUT a GPU: the product's Generator / Discriminator modules run on the emulated C ABI (tests/abi_emulator.py: numpy /
oracle restatements of the header's contract behind the product's own call sites) and are compared with the goldens the UNMODIFIED reference produced
at tensor-core-eligible widths (tests/golden/networks_wide.npz, oracle/make_golden.py networks_wide).  Same weights, latents, cameras, patch parameters
and injected noise as tests/test_gpu_networks_wide.py; the routing counters prove that the path taken is the fused one the GPU run takes.
Together with the GPU suite (kernels == contract) this closes  reference == host logic o contract  without a device."""
import importlib
import json
import os

import numpy as np
import pytest
import torch

import abi_emulator as emu
from conftest import ROOT
from oracle import cases
from util import maxrel, l2rel

pr = cases.grad_probe


@pytest.fixture(autouse=True)
def emulated(monkeypatch):
    tc = emu.install(monkeypatch)
    yield
    tc.invalidate_weight_cache()


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'networks_wide.npz'))


def _build():
    cfgm = importlib.import_module('3dgp_b200.config')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'networks_wide_meta.json')))
    kw = dict(meta['net_kwargs']); kw.pop('learn_camera_dist', None)
    cfg = cfgm.make_config(**kw, kd_weight=1.0)
    G, D = cfgm.build_networks(cfg, 'cpu', fp32_D=True)
    G.load_state_dict(cases.fill_state_dict({k: tuple(v) for k, v in meta['G_keys'].items()}, G.state_dict(), seed=100))
    D.load_state_dict(cases.fill_state_dict({k: tuple(v) for k, v in meta['D_keys'].items()}, D.state_dict(), seed=200))
    t = {k: torch.from_numpy(v) for k, v in cases.net_inputs(meta['net_kwargs']).items()}
    cam = dn.TensorGroup(angles=t['angles'], fov=t['fov'], radius=t['radius'], look_at=t['look_at'])
    pp = dict(scales=t['patch_scales'], offsets=t['patch_offsets'])
    return cfg, G, D, t, cam, pp, meta['net_kwargs']


def test_generator_training_forward_on_the_emulated_abi_matches_the_reference(gold):
    tc = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
    cfg, G, D, t, cam, pp, kw = _build()
    B = t['z'].shape[0]
    noises = [torch.from_numpy(n) for n in cases.layer_noises(kw, B)]
    G.train(); G.synthesis.nerf_noise_std = 0.0
    blocks = {}
    dec = G.synthesis.tri_plane_decoder
    hs = [getattr(dec, f'b{r}').register_forward_hook(lambda m, a, o, r=r: blocks.__setitem__(r, (o[0].detach(), o[1].detach()))) for r in dec.block_resolutions]
    s0 = dict(tc.stats)
    with torch.no_grad():
        ws = G.mapping(t['z'], t['c'])
        ro = dict(concat_depth=True, return_depth=True, u_coarse=t['u_coarse'], u_fine=t['u_fine'], depth_head_idx=torch.from_numpy(cases.depth_heads(B)), mlp_mode=0)
        out = G.synthesis(ws, cam, patch_params=pp, render_opts=ro, noise_mode='random', layer_noises=noises)
    for h in hs:
        h.remove()
    d = {k: tc.stats[k] - s0[k] for k in s0}
    assert d['fused'] == 14 and d['tc'] == 2 and d['aten'] == 4, d            # the routing tests/test_gpu_networks_wide.py asserts on the GPU
    assert maxrel(ws.numpy(), gold['G/ws']) < 1e-5
    errs = {}
    for r, (x, img) in blocks.items():
        errs[f'b{r}.x'] = maxrel(pr(x.contiguous().numpy()), gold[f'G/block/b{r}/x']); errs[f'b{r}.img'] = maxrel(pr(img.contiguous().numpy()), gold[f'G/block/b{r}/img'])
    errs['img'] = maxrel(out.img.numpy(), gold['G/train/img']); errs['depth'] = maxrel(out.depth.numpy(), gold['G/train/depth'])
    assert max(errs.values()) < 1e-4, errs                                     # float64 contractions on float32 storage: re-association level


def test_generator_eval_forward_on_the_emulated_abi_matches_the_reference(gold):
    """The G-inference path (metric_utils.py:303-319): eval mode, const noise, full-frame render at img_resolution, no patch."""
    tc = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
    cfg, G, D, t, cam, pp, kw = _build()
    G.eval()
    B = t['z'].shape[0]
    ue = cases.eval_variates(kw, B)
    ro = dict(concat_depth=True, return_depth=True, u_coarse=torch.from_numpy(ue['u_coarse']), u_fine=torch.from_numpy(ue['u_fine']), mlp_mode=0)
    s0 = dict(tc.stats)
    with torch.no_grad():
        ws = G.mapping(t['z'], t['c'])
        oe = G.synthesis(ws, cam, render_opts=ro, noise_mode='const')
    d = {k: tc.stats[k] - s0[k] for k in s0}
    assert d['fused'] == 14 and d['aten'] == 4, d
    assert maxrel(pr(oe.img.contiguous().numpy()), gold['G/eval/img']) < 1e-4
    assert maxrel(pr(oe.depth.contiguous().numpy()), gold['G/eval/depth']) < 1e-4


def test_discriminator_first_order_and_r1_on_the_emulated_abi_match_the_reference(gold):
    """Dmain (fused first-order nodes) and Dreg (the twice-differentiable composition, weight gradients off inside the inner pass: loss.py:238-253) with
    the routing counters of the GPU run."""
    tc = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
    cg = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix')
    layers = importlib.import_module('3dgp_b200.training.layers')
    cfg, G, D, t, cam, pp, kw = _build()
    D.train()
    B = t['z'].shape[0]
    names = cases.probe_params('D', 'wide')
    pars = dict(D.named_parameters())
    blocks = {}
    hs = [getattr(D, f'b{r}').register_forward_hook(lambda m, a, o, r=r: blocks.__setitem__(r, o.detach())) for r in D.block_resolutions]
    img = torch.from_numpy(gold['G/train/img']).requires_grad_(True)
    s0 = dict(tc.stats)
    logits, feats = D(img, t['c'], patch_params=pp, camera_angles=t['angles'], predict_feat=True)
    d = {k: tc.stats[k] - s0[k] for k in s0}
    for h in hs:
        h.remove()
    assert d['fused'] == 9 and d['tc'] == 6 and d['aten'] == 2, d
    ferr = {f'b{r}': maxrel(pr(x.contiguous().numpy()), gold[f'D/block/b{r}']) for r, x in blocks.items()}
    ferr['logits'] = maxrel(logits.detach().numpy(), gold['D/logits']); ferr['feats'] = maxrel(feats.detach().numpy(), gold['D/feats'])
    assert max(ferr.values()) < 1e-4, ferr
    embs = torch.from_numpy(cases.cotangent((B, kw['embedding_dim']), 31))
    loss1 = torch.nn.functional.softplus(-logits).mean() + (feats - embs).norm(dim=1).mean()
    assert abs(loss1.item() - float(gold['D/loss1'][0])) < 1e-4 * abs(float(gold['D/loss1'][0]))
    gs = torch.autograd.grad(loss1, [img] + [pars[n] for n in names])
    errs = {n: l2rel(pr(gr.contiguous().numpy()), gold['D/grad1/' + n]) for n, gr in zip(names, gs[1:])}
    errs['img'] = l2rel(gs[0].numpy(), gold['D/grad1/img'])
    assert max(errs.values()) < 5e-4, errs            # the emulated backward kernels hand bf16 (hi, lo) pairs on: ~2^-16 per product
    img = torch.from_numpy(gold['G/train/img']).requires_grad_(True)
    s0 = dict(tc.stats)
    with layers.first_order_only(False):
        logits, feats = D(img, t['c'], patch_params=pp, camera_angles=t['angles'], predict_feat=True)
    d = {k: tc.stats[k] - s0[k] for k in s0}
    assert d['fused'] == 0 and d['tc'] == 15 and d['aten'] == 2, d
    with cg.no_weight_gradients():
        r1 = torch.autograd.grad([logits.sum()], [img], create_graph=True)[0]
    loss = torch.nn.functional.softplus(-logits).mean() + r1.square().sum([1, 2, 3]).mean() * 0.5
    gs = torch.autograd.grad(loss, [pars[n] for n in names])
    errs = {n: l2rel(pr(gr.contiguous().numpy()), gold['D/grad/' + n]) for n, gr in zip(names, gs)}
    errs['r1(img)'] = l2rel(r1.detach().numpy(), gold['D/r1_grads'])
    assert max(errs.values()) < 5e-4, errs


def test_generator_loss_gradients_on_the_emulated_abi_match_the_reference(gold):
    """Gmain: softplus(-D(G(z))) through the fused D nodes, the ray-march wrapper's backward and the fused decoder nodes, against the reference's autograd."""
    tc = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
    cfg, G, D, t, cam, pp, kw = _build()
    B = t['z'].shape[0]
    noises = [torch.from_numpy(n) for n in cases.layer_noises(kw, B)]
    G.train(); G.synthesis.nerf_noise_std = 0.0
    D.train(); D.requires_grad_(False)
    s0 = dict(tc.stats)
    ws = G.mapping(t['z'], t['c'])
    ro = dict(concat_depth=True, return_depth=True, u_coarse=t['u_coarse'], u_fine=t['u_fine'], depth_head_idx=torch.from_numpy(cases.depth_heads(B)), mlp_mode=0)
    out = G.synthesis(ws, cam, patch_params=pp, render_opts=ro, noise_mode='random', layer_noises=noises)
    logits, _ = D(out.img, t['c'], patch_params=pp, camera_angles=t['angles'])
    d = {k: tc.stats[k] - s0[k] for k in s0}
    assert d['fused'] == 14 + 9 and d['tc'] == 2 + 6 and d['aten'] == 4 + 2, d
    loss = torch.nn.functional.softplus(-logits).mean()
    assert abs(loss.item() - float(gold['G/loss'][0])) < 1e-4 * max(1.0, abs(float(gold['G/loss'][0])))
    names = cases.probe_params('G', 'wide')
    pars = dict(G.named_parameters())
    gs = torch.autograd.grad(loss, [pars[n] for n in names])
    errs = {n: l2rel(pr(gr.contiguous().numpy()), gold['G/grad/' + n]) for n, gr in zip(names, gs)}
    assert max(errs.values()) < 1e-3, errs            # measured 4e-5 .. 1.2e-4, noise_strength 5.6e-4 (bf16 (hi, lo) pairs handed on by the emulated backward)


def test_one_training_iteration_of_the_real_modules_on_the_emulated_abi():
    """training/step.py::Trainer.step on CPU modules (its torch.optim path) with the product's loss (Gmain with the camera-adaptor regularisers, Dmain
    with the distillation term, lazy R1): every phase runs through the host logic the GPU step uses; stats finite, parameters and G_ema move."""
    cfgm = importlib.import_module('3dgp_b200.config')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    lossm = importlib.import_module('3dgp_b200.training.loss')
    stepm = importlib.import_module('3dgp_b200.training.step')
    meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'networks_meta.json')))
    kw = dict(meta['net_kwargs']); kw.pop('learn_camera_dist', None)
    cfg = cfgm.make_config(**kw, kd_weight=1.0, learn_camera_dist=True, batch_size=4)
    torch.manual_seed(0); np.random.seed(0)
    G, D = cfgm.build_networks(cfg, 'cpu', fp32_D=True)
    t = {k: torch.from_numpy(v) for k, v in cases.net_inputs(meta['net_kwargs']).items()}
    B = t['z'].shape[0]
    loss = lossm.StyleGAN2Loss(cfg, 'cpu', G, D, r1_gamma=1.0)
    loss.progressive_update(5000)
    tr = stepm.Trainer(G, D, loss, cfg, D_reg_interval=16, batch_size=B)
    assert not tr.flat                                                          # CPU modules: the torch.optim branch of the trainer
    res = kw['img_resolution']
    real = dn.EasyDict(img=torch.rand(B, 3, res, res) * 2 - 1, depth=torch.rand(B, 1, res, res) * 2 - 1, c=t['c'], embs=torch.randn(B, kw['embedding_dim']),
                       camera_angles=t['angles'])
    gen = dn.EasyDict(z=t['z'], c=t['c'], camera_params=dn.TensorGroup(angles=t['angles'], fov=t['fov'], radius=t['radius'], look_at=t['look_at']))
    w0 = G.synthesis.tri_plane_decoder.b8.conv0.weight.detach().clone(); d0 = D.b16.conv0.weight.detach().clone()
    e0 = tr.G_ema.synthesis.tri_plane_decoder.b8.conv0.weight.detach().clone()
    stats = tr.step(real, gen)
    assert all(torch.isfinite(torch.as_tensor(v)).all() for v in stats.values()), stats
    assert 'Loss/D/r1_penalty' in stats and 'Loss/camera_dist/emd_loss' in stats and 'Loss/camera_dist/force_mean' in stats
    assert not torch.equal(w0, G.synthesis.tri_plane_decoder.b8.conv0.weight) and not torch.equal(d0, D.b16.conv0.weight)
    assert not torch.equal(e0, tr.G_ema.synthesis.tri_plane_decoder.b8.conv0.weight)
    stats = tr.step(real, gen)
    assert 'Loss/D/r1_penalty' not in stats and tr.it == 2 and tr.cur_nimg == 2 * B


def test_snapshot_generator_to_uint8_images_on_the_emulated_abi():
    """Eval-side chain on CPU: reference-format snapshot pickle -> legacy.load_network_pkl -> inference.generate_uint8 (camera adaptor, G_ema, uint8
    conversion) == the module called directly + the reference expression of metric_utils.py:313."""
    lg = importlib.import_module('3dgp_b200.legacy')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    inf = importlib.import_module('3dgp_b200.training.inference')
    Ge = lg.load_network_pkl(os.path.join(ROOT, 'tests', 'golden', 'snapshot_small.pkl.gz'), device='cpu', names=('G_ema',))['G_ema']
    kw = cases.net_kwargs('small')
    t = {k: torch.from_numpy(v) for k, v in cases.net_inputs(kw).items()}
    cam = dn.TensorGroup(angles=t['angles'], fov=t['fov'], radius=t['radius'], look_at=t['look_at'])
    Ge.synthesis.renderer.launch_counter = 0
    img = inf.generate_uint8(Ge, t['z'], t['c'], cam, noise_mode='const')
    B = t['z'].shape[0]
    assert img.dtype == torch.uint8 and tuple(img.shape) == (B, 3, kw['img_resolution'], kw['img_resolution'])
    Ge.synthesis.renderer.launch_counter = 0
    with torch.no_grad():
        ref = Ge(z=t['z'], c=t['c'], camera_params=cam, camera_angles_cond=cam.angles, noise_mode='const')
    ref = ref if torch.is_tensor(ref) else ref.img
    assert torch.equal(img, (ref[:, :3] * 127.5 + 128).clamp(0, 255).to(torch.uint8))
    for cl in (False, True):          # NCHW and channels-last inputs of the conversion, 4-channel input -> first 3 channels
        x = torch.randn(2, 4, 8, 12) * 1.5
        xin = x.contiguous(memory_format=torch.channels_last) if cl else x
        assert torch.equal(inf.to_uint8(xin, channels=3), (x[:, :3] * 127.5 + 128).clamp(0, 255).to(torch.uint8))


# ---------------------------------------------------------------------------------------------------------------------------------
# The SMALL networks (32-channel layers: below the tensor-core kernels' channel granularity) take the UNFUSED composition -- x * styles ->
# conv2d_resample (stride-2 transposed conv + FIR for the up-sampling layers, FIR + stride-2 conv for the down-sampling ones) -> fma -> bias_act,
# the grouped-conv inference form, ATen convolutions -- i.e. the other half of the module code.  Same goldens as tests/test_gpu_networks.py.

def _build_small():
    cfgm = importlib.import_module('3dgp_b200.config')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'networks_meta.json')))
    kw = dict(meta['net_kwargs']); kw.pop('learn_camera_dist', None)
    cfg = cfgm.make_config(**kw, kd_weight=1.0)
    G, D = cfgm.build_networks(cfg, 'cpu', fp32_D=True)
    G.load_state_dict(cases.fill_state_dict({k: tuple(v) for k, v in meta['G_keys'].items()}, G.state_dict(), seed=100))
    D.load_state_dict(cases.fill_state_dict({k: tuple(v) for k, v in meta['D_keys'].items()}, D.state_dict(), seed=200))
    t = {k: torch.from_numpy(v) for k, v in cases.net_inputs(meta['net_kwargs']).items()}
    cam = dn.TensorGroup(angles=t['angles'], fov=t['fov'], radius=t['radius'], look_at=t['look_at'])
    return cfg, G, D, t, cam, dict(scales=t['patch_scales'], offsets=t['patch_offsets']), meta['net_kwargs']


def test_small_networks_unfused_composition_on_the_emulated_abi_match_the_reference():
    tc = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
    cg = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix')
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'networks.npz'))
    cfg, G, D, t, cam, pp, kw = _build_small()
    B = t['z'].shape[0]
    noises = [torch.from_numpy(n) for n in cases.layer_noises(kw, B)]
    G.train(); G.synthesis.nerf_noise_std = 0.0
    s0 = dict(tc.stats)
    ws = G.mapping(t['z'], t['c'])
    ro = dict(concat_depth=True, return_depth=True, u_coarse=t['u_coarse'], u_fine=t['u_fine'], depth_head_idx=torch.from_numpy(cases.depth_heads(B)), mlp_mode=0)
    out = G.synthesis(ws, cam, patch_params=pp, render_opts=ro, noise_mode='random', layer_noises=noises)
    assert tc.stats['fused'] == s0['fused'] and tc.stats['tc'] == s0['tc'], 'no layer of the small network is wide enough for the tensor-core kernels'
    assert maxrel(ws.detach().numpy(), gold['G/ws']) < 1e-5
    assert maxrel(out.img.detach().numpy(), gold['G/train/img']) < 1e-4 and maxrel(out.depth.detach().numpy(), gold['G/train/depth']) < 1e-4
    # Gmain gradients through D
    D.train()
    logits, _ = D(out.img, t['c'], patch_params=pp, camera_angles=t['angles'])
    loss = torch.nn.functional.softplus(-logits).mean()
    assert abs(loss.item() - float(gold['G/loss'][0])) < 1e-4 * max(1.0, abs(float(gold['G/loss'][0])))
    names = cases.probe_params('G'); pars = dict(G.named_parameters())
    for n, gr in zip(names, torch.autograd.grad(loss, [pars[n] for n in names])):
        assert l2rel(gr.numpy(), gold['G/grad/' + n]) < 5e-4, n
    # eval: grouped-conv (fused_modconv) inference form, const noise, full frame
    G.eval()
    ue = cases.eval_variates(kw, B)
    with torch.no_grad():
        oe = G.synthesis(ws.detach(), cam, render_opts=dict(concat_depth=True, return_depth=True, u_coarse=torch.from_numpy(ue['u_coarse']),
                                                            u_fine=torch.from_numpy(ue['u_fine']), mlp_mode=0), noise_mode='const')
    assert maxrel(oe.img.numpy(), gold['G/eval/img']) < 1e-4 and maxrel(oe.depth.numpy(), gold['G/eval/depth']) < 1e-4
    # D forward + R1 double backward
    img = torch.from_numpy(gold['G/train/img']).requires_grad_(True)
    logits, feats = D(img, t['c'], patch_params=pp, camera_angles=t['angles'], predict_feat=True)
    assert maxrel(logits.detach().numpy(), gold['D/logits']) < 1e-4 and maxrel(feats.detach().numpy(), gold['D/feats']) < 1e-4
    with cg.no_weight_gradients():
        r1 = torch.autograd.grad([logits.sum()], [img], create_graph=True)[0]
    assert l2rel(r1.detach().numpy(), gold['D/r1_grads']) < 1e-4
    loss = torch.nn.functional.softplus(-logits).mean() + r1.square().sum([1, 2, 3]).mean() * 0.5
    names = cases.probe_params('D'); pars = dict(D.named_parameters())
    for n, gr in zip(names, torch.autograd.grad(loss, [pars[n] for n in names])):
        assert l2rel(gr.numpy(), gold['D/grad/' + n]) < 5e-4, n
