"""GPU parity suite for the module surface (Generator / Discriminator / loss step) against goldens produced by the
unmodified reference on the reduced-width 3DGP config (oracle/cases.py:small_net_kwargs): identical weights (seeded
state dicts), latents, cameras, patch params and INJECTED layer / renderer noise."""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import cases
from util import maxrel, l2rel

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _build():
    cfgm = importlib.import_module('3dgp_b200.config')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'networks_meta.json')))
    kw = dict(meta['net_kwargs']); kw.pop('learn_camera_dist', None)
    cfg = cfgm.make_config(**kw, kd_weight=1.0)
    G, D = cfgm.build_networks(cfg, 'cuda', fp32_D=True)
    sdG = cases.fill_state_dict({k: tuple(v) for k, v in meta['G_keys'].items()}, G.state_dict(), seed=100)
    sdD = cases.fill_state_dict({k: tuple(v) for k, v in meta['D_keys'].items()}, D.state_dict(), seed=200)
    G.load_state_dict(sdG); D.load_state_dict(sdD)
    inp = cases.net_inputs(meta['net_kwargs'])
    t = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    cam = dn.TensorGroup(angles=t['angles'], fov=t['fov'], radius=t['radius'], look_at=t['look_at'])
    pp = dict(scales=t['patch_scales'], offsets=t['patch_offsets'])
    return cfg, G, D, t, cam, pp, meta['net_kwargs']


def _train_forward(G, t, cam, pp, kw):
    B = t['z'].shape[0]
    noises = [torch.from_numpy(n).cuda() for n in cases.layer_noises(kw, B)]
    G.train()
    G.synthesis.nerf_noise_std = 0.0
    ws = G.mapping(t['z'], t['c'])
    ro = dict(concat_depth=True, return_depth=True, u_coarse=t['u_coarse'], u_fine=t['u_fine'], depth_head_idx=torch.from_numpy(cases.depth_heads(B)))
    out = G.synthesis(ws, cam, patch_params=pp, render_opts=ro, noise_mode='random', layer_noises=noises)
    return ws, out, noises


def test_generator_train_and_eval_vs_golden(golden):
    cfg, G, D, t, cam, pp, kw = _build()
    g = golden('networks')
    ws, out, noises = _train_forward(G, t, cam, pp, kw)
    assert maxrel(ws.detach().cpu().numpy(), g['G/ws']) < 1e-5
    planes = G.synthesis.tri_plane_decoder(ws, noise_mode='random', layer_noises=noises, fused_modconv=False)
    assert maxrel(planes.detach().contiguous().flatten()[::31].cpu().numpy(), g['G/train/planes_probe']) < TOL
    assert maxrel(out.img.detach().cpu().numpy(), g['G/train/img']) < TOL
    assert maxrel(out.depth.detach().cpu().numpy(), g['G/train/depth']) < TOL
    # eval: fused modconv (grouped conv), const noise, full-frame render at img_resolution
    G.eval()
    B = t['z'].shape[0]
    ue = cases.eval_variates(kw, B)
    ro = dict(concat_depth=True, return_depth=True, u_coarse=torch.from_numpy(ue['u_coarse']).cuda(), u_fine=torch.from_numpy(ue['u_fine']).cuda())
    with torch.no_grad():
        oe = G.synthesis(ws, cam, render_opts=ro, noise_mode='const')
    assert maxrel(oe.img.cpu().numpy(), g['G/eval/img']) < TOL
    assert maxrel(oe.depth.cpu().numpy(), g['G/eval/depth']) < TOL


def test_discriminator_and_r1_double_backward_vs_golden(golden):
    cfg, G, D, t, cam, pp, kw = _build()
    cg = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix')
    g = golden('networks')
    D.train()
    img = torch.from_numpy(g['G/train/img']).cuda().requires_grad_(True)
    logits, feats = D(img, t['c'], patch_params=pp, camera_angles=t['angles'], predict_feat=True)
    assert maxrel(logits.detach().cpu().numpy(), g['D/logits']) < TOL
    assert maxrel(feats.detach().cpu().numpy(), g['D/feats']) < TOL
    with cg.no_weight_gradients():
        r1 = torch.autograd.grad([logits.sum()], [img], create_graph=True)[0]
    assert l2rel(r1.detach().cpu().numpy(), g['D/r1_grads']) < TOL
    loss = torch.nn.functional.softplus(-logits).mean() + r1.square().sum([1, 2, 3]).mean() * 0.5
    names = cases.probe_params('D')
    pars = dict(D.named_parameters())
    gs = torch.autograd.grad(loss, [pars[n] for n in names])
    for n, gr in zip(names, gs):
        assert l2rel(gr.cpu().numpy(), g['D/grad/' + n]) < 2e-3, n


def test_generator_loss_gradients_vs_golden(golden):
    cfg, G, D, t, cam, pp, kw = _build()
    g = golden('networks')
    D.train()
    ws, out, _ = _train_forward(G, t, cam, pp, kw)
    logits, _ = D(out.img, t['c'], patch_params=pp, camera_angles=t['angles'])
    loss = torch.nn.functional.softplus(-logits).mean()
    assert abs(loss.item() - float(g['G/loss'][0])) < 1e-3 * max(1.0, abs(float(g['G/loss'][0])))
    names = cases.probe_params('G')
    pars = dict(G.named_parameters())
    gs = torch.autograd.grad(loss, [pars[n] for n in names])
    for n, gr in zip(names, gs):
        assert l2rel(gr.cpu().numpy(), g['G/grad/' + n]) < 3e-3, n


def test_training_step_runs_and_updates():
    """One full Gmain + Dmain + Dreg iteration on the small config: finite stats, parameters move, G_ema tracks G."""
    cfg, G, D, t, cam, pp, kw = _build()
    dn = importlib.import_module('3dgp_b200.dnnlib')
    lossm = importlib.import_module('3dgp_b200.training.loss')
    stepm = importlib.import_module('3dgp_b200.training.step')
    B = t['z'].shape[0]
    loss = lossm.StyleGAN2Loss(cfg, 'cuda', G, D, r1_gamma=1.0)
    tr = stepm.Trainer(G, D, loss, cfg, D_reg_interval=16)
    torch.manual_seed(0); np.random.seed(0)
    real = dn.EasyDict(img=torch.rand(B, 3, 64, 64, device='cuda') * 2 - 1, depth=torch.rand(B, 1, 64, 64, device='cuda') * 2 - 1, c=t['c'],
                       embs=torch.randn(B, kw['embedding_dim'], device='cuda'), camera_angles=t['angles'])
    gen = dn.EasyDict(z=t['z'], c=t['c'], camera_params=cam)
    w0 = G.synthesis.tri_plane_decoder.b8.conv0.weight.detach().clone(); d0 = D.b16.conv0.weight.detach().clone()
    stats = tr.step(real, gen)
    assert all(torch.isfinite(v).all() for v in stats.values()) and 'Loss/D/r1_penalty' in stats
    assert not torch.equal(w0, G.synthesis.tri_plane_decoder.b8.conv0.weight) and not torch.equal(d0, D.b16.conv0.weight)
    stats = tr.step(real, gen)
    assert 'Loss/D/r1_penalty' not in stats     # lazy regularisation: Dreg only every 16th iteration


def test_camera_adaptor_vs_golden(golden):
    """Learned camera distribution (networks_camera_adaptor.py): the reference's weights, prior cameras, z, c -> posterior camera and gradients."""
    ca_mod = importlib.import_module('3dgp_b200.training.networks_camera_adaptor')
    cfgm = importlib.import_module('3dgp_b200.config')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    g = golden('camera_adaptor')
    cfg = cfgm.make_config(learn_camera_dist=True, z_dim=16, c_dim=6).model.generator.camera_adaptor
    cfg.hid_dim = 32; cfg.embed_dim = 8
    ca = ca_mod.CameraAdaptor(cfg).cuda()
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('sd/')}
    assert set(sd) == set(ca.state_dict())                      # same module / parameter names as the reference
    ca.load_state_dict(sd)
    cu = lambda k: torch.from_numpy(g[k]).cuda()
    cam = dn.TensorGroup(angles=cu('in/angles'), fov=cu('in/fov'), radius=cu('in/radius'), look_at=cu('in/look_at'))
    z = cu('in/z').requires_grad_(True)
    out = ca(cam, z, cu('in/c'))
    raw = ca.unroll_camera_params(out)
    assert maxrel(raw.detach().cpu().numpy(), g['out/raw']) < 1e-5
    names = [k[5:] for k in g.files if k.startswith('grad/') and k != 'grad/z']
    pars = dict(ca.named_parameters())
    grads = torch.autograd.grad((raw * cu('in/cot')).sum(), [z] + [pars[n] for n in names])
    assert maxrel(grads[0].cpu().numpy(), g['grad/z']) < 1e-4
    for n, gr in zip(names, grads[1:]):
        assert maxrel(gr.cpu().numpy(), g['grad/' + n]) < 1e-4, n


def test_training_step_with_learned_camera_distribution():
    """training.learn_camera_dist=true on the small config: the camera adaptor sits in front of the renderer, its weights receive gradient
    through d(ray_o), d(ray_d) of the fused ray-march backward plus the EMD / force-mean regularisers, and move."""
    cfgm = importlib.import_module('3dgp_b200.config')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    lossm = importlib.import_module('3dgp_b200.training.loss')
    stepm = importlib.import_module('3dgp_b200.training.step')
    meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'networks_meta.json')))
    kw = dict(meta['net_kwargs']); kw.pop('learn_camera_dist', None)
    cfg = cfgm.make_config(**kw, kd_weight=1.0, learn_camera_dist=True)
    torch.manual_seed(0); np.random.seed(0)
    G, D = cfgm.build_networks(cfg, 'cuda', fp32_D=True)
    assert G.synthesis.camera_adaptor is not None
    inp = cases.net_inputs(meta['net_kwargs'])
    t = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    cam = dn.TensorGroup(angles=t['angles'], fov=t['fov'], radius=t['radius'], look_at=t['look_at'])
    B = t['z'].shape[0]
    loss = lossm.StyleGAN2Loss(cfg, 'cuda', G, D, r1_gamma=1.0)
    loss.progressive_update(5000)                      # EMD weight ramps in from 0 (loss.py:64-65)
    assert loss.emd_multiplier > 0
    # gradient reaches the adaptor through the renderer alone
    G.requires_grad_(True); D.requires_grad_(False)
    out, pp = loss.run_G(t['z'], t['c'], cam)
    ca = G.synthesis.camera_adaptor
    gw = torch.autograd.grad(out.img.square().mean(), [ca.origin_adaptor.main[0].weight, ca.look_at_adaptor.main[1].weight], allow_unused=False)
    assert all(torch.isfinite(x).all() and x.abs().max() > 0 for x in gw)
    tr = stepm.Trainer(G, D, loss, cfg, D_reg_interval=16)
    real = dn.EasyDict(img=torch.rand(B, 3, 64, 64, device='cuda') * 2 - 1, depth=torch.rand(B, 1, 64, 64, device='cuda') * 2 - 1, c=t['c'],
                       embs=torch.randn(B, kw['embedding_dim'], device='cuda'), camera_angles=t['angles'])
    gen = dn.EasyDict(z=t['z'], c=t['c'], camera_params=cam)
    w0 = ca.look_at_adaptor.main[0].weight.detach().clone()
    stats = tr.step(real, gen)
    assert all(torch.isfinite(v).all() for v in stats.values())
    assert 'Loss/camera_dist/emd_loss' in stats and 'Loss/camera_dist/force_mean' in stats
    assert not torch.equal(w0, ca.look_at_adaptor.main[0].weight)


def test_training_iterations_on_a_training_set_read_from_disk(tmp_path):
    """training/training_loop.py::training_iterations: the reference loop's per-iteration data path (training_loop.py:296-366) end to end on the small
    config -- PNG / depth / label / camera-angle / embedding files -> worker threads -> pinned ring -> H2D prefetch -> device normalisation ->
    generator-side batch drawn from the training set -> Trainer.step (Gmain, Dmain with the distillation term, lazy R1)."""
    from util import write_training_set
    cfgm = importlib.import_module('3dgp_b200.config')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    lossm = importlib.import_module('3dgp_b200.training.loss')
    stepm = importlib.import_module('3dgp_b200.training.step')
    dsmod = importlib.import_module('3dgp_b200.training.dataset')
    tl = importlib.import_module('3dgp_b200.training.training_loop')
    meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'networks_meta.json')))
    kw = dict(meta['net_kwargs']); kw.pop('learn_camera_dist', None)
    cfg = cfgm.make_config(**kw, kd_weight=1.0, batch_size=4)
    torch.manual_seed(0); np.random.seed(0)
    G, D = cfgm.build_networks(cfg, 'cuda', fp32_D=True)
    root = str(tmp_path / 'train')
    extra = write_training_set(root, n=12, res=kw['img_resolution'], c_dim=kw['c_dim'], depth=True, emb_dim=kw['embedding_dim'])
    extra.pop('_embeddings'); extra.pop('_rows')
    dcfg = dn.EasyDict.init_recursively(dict(c_dim=kw['c_dim'], mirror=True, camera=cfg.camera, **extra))
    ds = dsmod.ImageFolderDataset(path=root, resolution=kw['img_resolution'], use_depth=True, cfg=dcfg)
    assert len(ds) == 24 and ds.label_dim == kw['c_dim']
    loss = lossm.StyleGAN2Loss(cfg, 'cuda', G, D, r1_gamma=1.0)
    tr = stepm.Trainer(G, D, loss, cfg, D_reg_interval=16)
    w0 = G.synthesis.tri_plane_decoder.b8.conv0.weight.detach().clone(); d0 = D.b16.conv0.weight.detach().clone()
    all_stats = list(tl.training_iterations(tr, ds, 'cuda', 3, batch=4, workers=3))
    assert len(all_stats) == 3 and tr.it == 3 and tr.cur_nimg == 12
    for st in all_stats:
        assert all(torch.isfinite(v).all() for v in st.values())
    assert 'Loss/D/r1_penalty' in all_stats[0] and 'Loss/D/r1_penalty' not in all_stats[1]
    assert not torch.equal(w0, G.synthesis.tri_plane_decoder.b8.conv0.weight) and not torch.equal(d0, D.b16.conv0.weight)
