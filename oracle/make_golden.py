"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_harness.py) on seeded inputs, and pins oracle/restated.py against it.

    python oracle/make_golden.py            # regenerate fixtures + print oracle-vs-reference deviations

Inputs are produced by `oracle/cases.py` from numpy RandomState seeds, so fixtures only store OUTPUTS (plus the
state-dict key/shape lists of the small networks); tests regenerate the identical inputs on any machine.
Runs only where /root/reference exists (the build container); the GPU box consumes the committed fixtures.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases, ref_harness as rh, restated as R  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def maxrel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def gen_upfirdn2d(ns, report):
    out = {}
    for name, kw in cases.upfirdn2d_cases():
        x, f = cases.upfirdn2d_inputs(name, kw)
        y_ref = ns.upfirdn2d._upfirdn2d_ref(torch.from_numpy(x), None if f is None else torch.from_numpy(f), up=kw['up'], down=kw['down'],
                                            padding=kw['padding'], flip_filter=kw['flip_filter'], gain=kw['gain']).numpy()
        y_or = R.upfirdn2d(x, f, up=kw['up'], down=kw['down'], padding=kw['padding'], flip_filter=kw['flip_filter'], gain=kw['gain'])
        assert y_or.shape == y_ref.shape, (name, y_or.shape, y_ref.shape)
        report[f'upfirdn2d/{name}'] = maxrel(y_or, y_ref)
        if kw.get('integer', False):
            assert np.array_equal(y_or, y_ref), f'{name}: integer-valued case must be bit-exact'
        out[name] = y_ref
    np.savez_compressed(os.path.join(GOLD, 'upfirdn2d.npz'), **out)


def gen_bias_act(ns, report):
    out = {}
    for name, kw in cases.bias_act_cases():
        x, b = cases.bias_act_inputs(name, kw)
        xt = torch.from_numpy(x).requires_grad_(True)
        bt = torch.from_numpy(b).requires_grad_(True) if b is not None else None
        y = ns.bias_act._bias_act_ref(xt, bt, dim=kw['dim'], act=kw['act'], alpha=kw.get('alpha'), gain=kw.get('gain'), clamp=kw.get('clamp'))
        y_or = R.bias_act(x, b, dim=kw['dim'], act=kw['act'], alpha=kw.get('alpha'), gain=kw.get('gain'), clamp=kw.get('clamp'))
        report[f'bias_act/{name}'] = maxrel(y_or, y.detach().numpy())
        out[name + '/y'] = y.detach().numpy()
        # first-order grads for dy = cos(arange)
        dy = torch.from_numpy(cases.cotangent(y.shape, 11))
        g = torch.autograd.grad(y, [xt] + ([bt] if bt is not None else []), dy, create_graph=True)
        out[name + '/dx'] = g[0].detach().numpy()
        if bt is not None:
            out[name + '/db'] = g[1].detach().numpy()
        # second order: d/dx of <dx, v>
        v = torch.from_numpy(cases.cotangent(y.shape, 12))
        if g[0].requires_grad:
            g2 = torch.autograd.grad(g[0], xt, v, allow_unused=True)[0]
            out[name + '/d2x'] = (g2 if g2 is not None else torch.zeros_like(xt)).detach().numpy()
    np.savez_compressed(os.path.join(GOLD, 'bias_act.npz'), **out)


def gen_filtered_lrelu(ns, report):
    out = {}
    for name, kw in cases.filtered_lrelu_cases():
        x, fu, fd, b = cases.filtered_lrelu_inputs(name, kw)
        xt = torch.from_numpy(x).requires_grad_(True)
        bt = torch.from_numpy(b).requires_grad_(True)
        y = ns.filtered_lrelu._filtered_lrelu_ref(xt, torch.from_numpy(fu), torch.from_numpy(fd), bt, up=kw['up'], down=kw['down'],
                                                   padding=kw['padding'], gain=kw['gain'], slope=kw['slope'], clamp=kw['clamp'])
        y_or = R.filtered_lrelu(x, fu, fd, b, up=kw['up'], down=kw['down'], padding=kw['padding'], gain=kw['gain'], slope=kw['slope'], clamp=kw['clamp'])
        report[f'filtered_lrelu/{name}'] = maxrel(y_or, y.detach().numpy())
        dy = torch.from_numpy(cases.cotangent(y.shape, 13))
        gx, gb = torch.autograd.grad(y, [xt, bt], dy)
        out[name + '/y'] = y.detach().numpy(); out[name + '/dx'] = gx.numpy(); out[name + '/db'] = gb.numpy()
    np.savez_compressed(os.path.join(GOLD, 'filtered_lrelu.npz'), **out)


def gen_render(ns, report):
    out = {}
    ED = ns.dnnlib.EasyDict
    for name, kw in cases.render_cases():
        inp = cases.render_inputs(name, kw)
        t = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in inp.items()}
        Gc, _, _ = rh.make_cfg(num_ray_steps=kw['N'], tri_res=kw['P'])
        cfg = ED.init_recursively(Gc)
        mlp = ns.networks_epigraf.TriPlaneMLP(cfg, out_dim=3)
        with torch.no_grad():
            mlp.model[0].weight.copy_(t['w1']); mlp.model[0].bias.copy_(t['b1'])
            mlp.model[1].weight.copy_(t['w2']); mlp.model[1].bias.copy_(t['b2'])
        renderer = ns.tri_plane_renderer.ImportanceRenderer('classical')
        renderer.train(kw.get('training', True))
        opts = ED(box_size=kw['box_half'] * 2, num_proposal_steps=kw['N'], clamp_mode=kw.get('clamp_mode', 'softplus'),
                  use_inf_depth=kw.get('use_inf_depth', True), ray_start=kw['ray_start'], ray_end=kw['ray_end'], num_fine_steps=kw['N'],
                  density_noise=kw.get('noise_std', 0.0), last_back=kw.get('last_back', False), white_back=False,
                  max_batch_res=128, cut_quantile=0.0, density_bias=0.0)
        if kw.get('white_back_end_idx', 0):
            opts.white_back_end_idx = kw['white_back_end_idx']
        planes = t['planes'].clone().requires_grad_(True)
        ro = t['ray_o'].clone().requires_grad_(True); rd = t['ray_d'].clone().requires_grad_(True)
        B, Rr, N = kw['B'], kw['R'], kw['N']
        randn_like = []
        if kw.get('noise_std', 0.0) > 0:
            randn_like = [t['sn_coarse'].reshape(B, Rr * N, 1), t['sn_fine'].reshape(B, Rr * N, 1)]
        with rh.injected_rng(rand_like=[t['u_coarse'].reshape(B, Rr, N, 1)], rand=[t['u_fine'].reshape(B * Rr, N)], randn_like=randn_like):
            rgb, depth, wsum, tfin = renderer(planes, mlp, ro, rd, opts)
        # oracle
        o_rgb, o_depth, o_wsum, o_T = R.render(
            t['planes'], t['w1'], t['b1'], t['w2'], t['b2'], t['ray_o'], t['ray_d'], t['u_coarse'], t['u_fine'], kw['ray_start'], kw['ray_end'],
            kw['box_half'], N, sn_coarse=t.get('sn_coarse'), sn_fine=t.get('sn_fine'), noise_std=kw.get('noise_std', 0.0),
            use_inf_depth=kw.get('use_inf_depth', True), last_back=kw.get('last_back', False),
            white_back_end_idx=kw.get('white_back_end_idx', 0), clamp_mode=kw.get('clamp_mode', 'softplus'))
        report[f'render/{name}/rgb'] = maxrel(o_rgb, rgb.detach()); report[f'render/{name}/depth'] = maxrel(o_depth, depth.detach().squeeze(-1))
        report[f'render/{name}/wsum'] = maxrel(o_wsum, wsum.detach().squeeze(-1)); report[f'render/{name}/T'] = maxrel(o_T, tfin.detach())
        out[name + '/rgb'] = rgb.detach().numpy(); out[name + '/depth'] = depth.detach().squeeze(-1).numpy()
        out[name + '/wsum'] = wsum.detach().squeeze(-1).numpy(); out[name + '/tfinal'] = tfin.detach().numpy()
        # gradients for the backward kernel: L = <rgb, g_rgb> + <depth, g_depth>
        g_rgb = torch.from_numpy(cases.cotangent(rgb.shape, 21)); g_dep = torch.from_numpy(cases.cotangent(depth.shape, 22))
        params = [planes, mlp.model[0].weight, mlp.model[0].bias, mlp.model[1].weight, mlp.model[1].bias, ro, rd]
        grads = torch.autograd.grad([rgb, depth], params, [g_rgb, g_dep])
        for nm, g in zip(['g_planes', 'g_w1', 'g_b1', 'g_w2', 'g_b2', 'g_ray_o', 'g_ray_d'], grads):
            if nm == 'g_planes' and g.numel() > 400000:
                # large plane gradient: keep a strided probe + global statistics (the test recomputes the same reductions)
                out[name + '/g_planes_probe'] = g.flatten()[::97].numpy().copy()
                out[name + '/g_planes_sum'] = np.array([g.double().sum().item(), g.double().abs().sum().item(), g.double().square().sum().item()])
            else:
                out[name + '/' + nm] = g.numpy()
    np.savez_compressed(os.path.join(GOLD, 'render.npz'), **out)


def gen_networks(ns, report, variant='small'):
    """variant 'small': 32-channel networks (every conv below the tensor-core path's channel granularity -> ATen in the product);
    variant 'wide': cases.wide_net_kwargs() -- every hot conv eligible for the tcgen05 path; large tensors are stored as strided probes."""
    out = {}
    meta = {}
    ED = ns.dnnlib.EasyDict
    kw = cases.net_kwargs(variant)
    wide = variant != 'small'
    tag = 'networks' if not wide else 'networks_' + variant
    rp = (lambda k: k) if not wide else (lambda k: f'{variant}/{k}')
    gp_ = (lambda g: g.numpy()) if not wide else (lambda g: cases.grad_probe(g.numpy()))
    Gc, Dc, m = rh.make_cfg(**kw)
    G = rh.build_reference_G(Gc, m['img_resolution'], seed=0)
    D = rh.build_reference_D(Dc, m['patch_res'], use_depth=True, embedding_dim=m['embedding_dim'], seed=1, fp32=True)
    # deterministic, machine-independent weights
    sdG = cases.fill_state_dict({k: tuple(v.shape) for k, v in G.state_dict().items()}, G.state_dict(), seed=100)
    sdD = cases.fill_state_dict({k: tuple(v.shape) for k, v in D.state_dict().items()}, D.state_dict(), seed=200)
    G.load_state_dict(sdG); D.load_state_dict(sdD)
    meta['G_keys'] = {k: list(v.shape) for k, v in G.state_dict().items()}
    meta['D_keys'] = {k: list(v.shape) for k, v in D.state_dict().items()}
    meta['net_kwargs'] = kw
    inp = cases.net_inputs(kw)
    t = {k: torch.from_numpy(v) for k, v in inp.items()}
    B = t['z'].shape[0]
    cam = ns.dnnlib.TensorGroup(angles=t['angles'], fov=t['fov'], radius=t['radius'], look_at=t['look_at'])
    pp = dict(scales=t['patch_scales'], offsets=t['patch_offsets'])
    N = kw['num_ray_steps']; pr = kw['patch_res']; Rr = pr * pr

    # --- mapping
    ws = G.mapping(t['z'], t['c'])
    out['G/ws'] = ws.detach().numpy()
    ws_or = R.mapping_network(sdG, 'mapping.', t['z'], t['c'], G.num_ws)
    report[rp('G/mapping')] = maxrel(ws_or, ws.detach())

    # --- training-mode synthesis (fused_modconv=False, random layer noise injected, patch render, depth head fixed by np seed)
    G.train()
    layer_noise = [torch.from_numpy(n) for n in cases.layer_noises(kw, B)]
    head_idx = cases.depth_heads(B)
    np.random.seed(1234)
    orig_choice = np.random.choice
    np.random.choice = lambda *a, **k: head_idx.copy()
    hooks = []
    if wide:   # per-block activations (strided probes): localise a deviation to a block instead of to "the decoder"
        dec = G.synthesis.tri_plane_decoder
        for res in dec.block_resolutions:
            def hk(mod, args, outs, res=res):
                out[f'G/block/b{res}/x'] = cases.grad_probe(outs[0].detach().numpy()); out[f'G/block/b{res}/img'] = cases.grad_probe(outs[1].detach().numpy())
            hooks.append(getattr(dec, f'b{res}').register_forward_hook(hk))
    try:
        with rh.injected_rng(randn=[n.clone() for n in layer_noise], rand_like=[t['u_coarse'].reshape(B, Rr, N, 1)], rand=[t['u_fine'].reshape(B * Rr, N)]):
            G.synthesis.nerf_noise_std = 0.0
            o = G.synthesis(ws, cam, patch_params=pp, render_opts=dict(concat_depth=True, return_depth=True))
    finally:
        np.random.choice = orig_choice
        for h in hooks:
            h.remove()
    out['G/train/img'] = o.img.detach().numpy(); out['G/train/depth'] = o.depth.detach().numpy()
    # decoder planes via the reference decoder alone (same noise)
    with rh.injected_rng(randn=[n.clone() for n in layer_noise]):
        planes_ref = G.synthesis.tri_plane_decoder(ws, noise_mode='random', fused_modconv=False)
    out['G/train/planes_probe'] = planes_ref.detach().flatten()[::31].numpy().copy()
    out['G/train/planes_stats'] = np.array([planes_ref.double().sum().item(), planes_ref.double().abs().sum().item()])
    o_or = R.generator_synthesis(sdG, Gc, ws.detach(), t['angles'], t['fov'], t['radius'], t['look_at'], pr, t['patch_scales'], t['patch_offsets'],
                                 t['u_coarse'], t['u_fine'], noise_mode='random', noises=layer_noise, fused_modconv=False,
                                 depth_head_idx=torch.from_numpy(head_idx))
    report[rp('G/train/planes')] = maxrel(o_or['planes'], planes_ref.detach())
    report[rp('G/train/img')] = maxrel(o_or['img'], o.img.detach()); report[rp('G/train/depth')] = maxrel(o_or['depth'], o.depth.detach())

    # --- eval-mode synthesis (fused modconv, const noise, full-frame render at img_resolution)
    G.eval()
    res = kw['img_resolution']; Re = res * res
    ue = cases.eval_variates(kw, B)
    with rh.injected_rng(rand_like=[torch.from_numpy(ue['u_coarse']).reshape(B, Re, N, 1)], rand=[torch.from_numpy(ue['u_fine']).reshape(B * Re, N)]):
        oe = G.synthesis(ws, cam, noise_mode='const', render_opts=dict(concat_depth=True, return_depth=True))
    out['G/eval/img'] = gp_(oe.img.detach()) if wide else oe.img.detach().numpy()
    out['G/eval/depth'] = gp_(oe.depth.detach()) if wide else oe.depth.detach().numpy()
    oe_or = R.generator_synthesis(sdG, Gc, ws.detach(), t['angles'], t['fov'], t['radius'], t['look_at'], res, None, None,
                                  torch.from_numpy(ue['u_coarse']), torch.from_numpy(ue['u_fine']), noise_mode='const', fused_modconv=True)
    report[rp('G/eval/img')] = maxrel(oe_or['img'], oe.img.detach())

    # --- discriminator on the training-mode fake patch + R1-style double backward
    D.train()
    img = o.img.detach().clone().requires_grad_(True)
    hooks = []
    if wide:
        for res in D.block_resolutions:
            def hk(mod, args, outs, res=res):
                out[f'D/block/b{res}'] = cases.grad_probe(outs.detach().numpy())
            hooks.append(getattr(D, f'b{res}').register_forward_hook(hk))
    logits, feats = D(img, t['c'], patch_params=pp, camera_angles=t['angles'], predict_feat=True)
    for h in hooks:
        h.remove()
    out['D/logits'] = logits.detach().numpy(); out['D/feats'] = feats.detach().numpy()
    lg_or, f_or = R.discriminator(sdD, o.img.detach(), t['c'], t['patch_scales'], t['patch_offsets'], D.block_resolutions,
                                  Dc['num_additional_start_blocks'], predict_feat=True)
    report[rp('D/logits')] = maxrel(lg_or, logits.detach()); report[rp('D/feats')] = maxrel(f_or, feats.detach())
    with ns.conv2d_gradfix.no_weight_gradients():
        r1 = torch.autograd.grad([logits.sum()], [img], create_graph=True)[0]
    out['D/r1_grads'] = r1.detach().numpy()
    pen = r1.square().sum([1, 2, 3])
    loss = torch.nn.functional.softplus(-logits).mean() + pen.mean() * 0.5
    names = cases.probe_params('D', variant)
    pars = dict(D.named_parameters())
    gs = torch.autograd.grad(loss, [pars[n] for n in names])
    for n, g in zip(names, gs):
        out['D/grad/' + n] = gp_(g)

    if wide:
        # --- first-order discriminator pass (Dmain on a given patch: adversarial + knowledge-distillation terms, loss.py:256-316): the phase the
        # fused first-order nodes of the product serve (R1 above needs the twice-differentiable composition)
        img1 = o.img.detach().clone().requires_grad_(True)
        lg1, f1 = D(img1, t['c'], patch_params=pp, camera_angles=t['angles'], predict_feat=True)
        embs = torch.from_numpy(cases.cotangent((B, kw['embedding_dim']), 31))
        loss1 = torch.nn.functional.softplus(-lg1).mean() + (f1 - embs).norm(dim=1).mean()
        gs1 = torch.autograd.grad(loss1, [img1] + [pars[n] for n in names])
        out['D/loss1'] = np.array([loss1.item()])
        out['D/grad1/img'] = gs1[0].numpy()
        for n, g in zip(names, gs1[1:]):
            out['D/grad1/' + n] = gp_(g)

    # --- generator loss gradient through D (Gmain): d softplus(-D(G(z))) / d params
    G.train()
    for p in G.parameters():
        p.grad = None
    np.random.choice = lambda *a, **k: head_idx.copy()
    try:
        with rh.injected_rng(randn=[n.clone() for n in layer_noise], rand_like=[t['u_coarse'].reshape(B, Rr, N, 1)], rand=[t['u_fine'].reshape(B * Rr, N)]):
            ws2 = G.mapping(t['z'], t['c'])
            o2 = G.synthesis(ws2, cam, patch_params=pp, render_opts=dict(concat_depth=True, return_depth=True))
    finally:
        np.random.choice = orig_choice
    lg2, _ = D(o2.img, t['c'], patch_params=pp, camera_angles=t['angles'])
    lossG = torch.nn.functional.softplus(-lg2).mean()
    namesG = cases.probe_params('G', variant)
    parsG = dict(G.named_parameters())
    gsG = torch.autograd.grad(lossG, [parsG[n] for n in namesG])
    out['G/loss'] = np.array([lossG.item()])
    for n, g in zip(namesG, gsG):
        out['G/grad/' + n] = gp_(g)

    np.savez_compressed(os.path.join(GOLD, tag + '.npz'), **out)
    with open(os.path.join(GOLD, tag + '_meta.json'), 'w') as f:
        json.dump(meta, f, indent=1, sort_keys=True)


def gen_camera_adaptor(ns, report):
    """Learned camera distribution (networks_camera_adaptor.py:54-134) of the UNMODIFIED reference on CPU: seeded weights (stored), prior
    cameras / z / c from numpy seeds; outputs + gradients of a fixed cotangent w.r.t. z and two weight tensors."""
    import importlib
    ca_mod = importlib.import_module('src.training.networks_camera_adaptor')
    G_cfg, _, _ = rh.make_cfg(learn_camera_dist=True, z_dim=16, c_dim=6)
    cfg = ns.dnnlib.EasyDict.init_recursively(G_cfg['camera_adaptor'])
    cfg.hid_dim = 32; cfg.embed_dim = 8
    torch.manual_seed(3)
    ca = ca_mod.CameraAdaptor(cfg)
    with torch.no_grad():                      # biases start at 0 in the reference; give them values so that they are exercised
        for n_, p_ in ca.named_parameters():
            if n_.endswith('bias'):
                p_.normal_(0, 1.0)
    rs = np.random.RandomState(7)
    B = 9
    cam = ns.dnnlib.TensorGroup(
        angles=torch.from_numpy(np.stack([rs.uniform(-1.5, 1.5, B), rs.uniform(0.8, 2.3, B), np.zeros(B)], 1).astype(np.float32)),
        fov=torch.from_numpy(rs.uniform(10, 45, B).astype(np.float32)), radius=torch.ones(B),
        look_at=torch.from_numpy(np.stack([rs.uniform(-3, 3, B), rs.uniform(0.1, 3.0, B), rs.uniform(0, 0.2, B)], 1).astype(np.float32)))
    z = torch.from_numpy(rs.randn(B, 16).astype(np.float32)).requires_grad_(True)
    c = torch.zeros(B, 6); c[torch.arange(B), torch.from_numpy(rs.randint(0, 6, B))] = 1.0
    out = ca(cam, z, c)
    raw = ca.unroll_camera_params(out)
    cot = torch.from_numpy(rs.randn(B, 8).astype(np.float32))
    names = ['origin_adaptor.main.0.weight', 'look_at_adaptor.project_z.weight', 'look_at_adaptor.main.1.bias']
    pars = dict(ca.named_parameters())
    grads = torch.autograd.grad((raw * cot).sum(), [z] + [pars[n_] for n_ in names])
    res = {'in/angles': cam.angles.numpy(), 'in/fov': cam.fov.numpy(), 'in/radius': cam.radius.numpy(), 'in/look_at': cam.look_at.numpy(),
           'in/z': z.detach().numpy(), 'in/c': c.numpy(), 'in/cot': cot.numpy(), 'out/raw': raw.detach().numpy(), 'grad/z': grads[0].numpy()}
    for n_, g_ in zip(names, grads[1:]):
        res['grad/' + n_] = g_.numpy()
    for k, v in ca.state_dict().items():
        res['sd/' + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, 'camera_adaptor.npz'), **res)


def gen_snapshot(ns, report):
    """A network snapshot exactly as src/training/training_loop.py:478-484 pickles it (persistent_class objects), for the small golden networks, with the
    embedded module SOURCES replaced by a placeholder (reference source code must not enter this repository; 3dgp_b200/legacy.py never reads them)."""
    import copy
    import pickle
    kw = cases.net_kwargs('small')
    Gc, Dc, m = rh.make_cfg(**kw)
    G = rh.build_reference_G(Gc, m['img_resolution'], seed=0)
    D = rh.build_reference_D(Dc, m['patch_res'], use_depth=True, embedding_dim=m['embedding_dim'], seed=1, fp32=True)
    import gzip
    for net, tag in ((G, 'G.'), (D, 'D.')):
        net.load_state_dict({k: torch.from_numpy(cases.snapshot_fill(tag + k, tuple(v.shape))).to(v.dtype) for k, v in net.state_dict().items()})
    G_ema = copy.deepcopy(G).eval()
    for net in (G, D, G_ema):
        for mod in net.modules():
            if hasattr(type(mod), '_orig_module_src'):
                type(mod)._orig_module_src = '# module source stripped from the fixture (oracle/make_golden.py::gen_snapshot)'
    data = dict(G=G, D=D, G_ema=G_ema, augment_pipe=None, training_set_kwargs=dict(path='synthetic', resolution=m['img_resolution'], use_labels=True))
    with gzip.open(os.path.join(GOLD, 'snapshot_small.pkl.gz'), 'wb') as f:
        pickle.dump(data, f)


def gen_reference_config(ns, report):
    """The experiment configuration the reference's launcher would save for the README's ImageNet-256 command (oracle/compose_config.py)."""
    from oracle import compose_config as cc
    cfg = cc.compose(os.path.join(rh.REF_ROOT, 'configs'), cc.README_IMAGENET_OVERRIDES)
    with open(os.path.join(GOLD, 'reference_experiment_config.json'), 'w') as f:
        json.dump(cfg, f, indent=1, sort_keys=True)


def main():
    assert rh.available(), 'reference not found'
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ns = rh.load()
    report = {}
    only = sys.argv[1:]
    gens = dict(upfirdn2d=gen_upfirdn2d, bias_act=gen_bias_act, filtered_lrelu=gen_filtered_lrelu, render=gen_render, networks=gen_networks,
                networks_wide=lambda ns_, rep_: gen_networks(ns_, rep_, 'wide'), camera_adaptor=gen_camera_adaptor,
                reference_config=gen_reference_config, snapshot=gen_snapshot)
    for name, fn in gens.items():
        if only and name not in only:
            continue
        fn(ns, report)
        print(f'[make_golden] {name} done', flush=True)
    worst = max(report.values()) if report else 0.0
    for k in sorted(report):
        print(f'  oracle-vs-reference {k:48s} max-rel {report[k]:.3e}')
    print(f'[make_golden] worst oracle-vs-reference deviation: {worst:.3e}')
    rp = os.path.join(GOLD, 'oracle_pin_report.json')
    old = {}
    if only and os.path.exists(rp):
        old = json.load(open(rp))
    old.update(report)
    with open(rp, 'w') as f:
        json.dump(old, f, indent=1, sort_keys=True)
    assert worst < 2e-4, 'oracle restatement deviates from the reference'


if __name__ == '__main__':
    main()
