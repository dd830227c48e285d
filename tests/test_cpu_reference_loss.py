"""Module-level overlay of INTEGRATION.md, executed from the reference's side: the UNMODIFIED reference training phases
(src/training/loss.py::StyleGAN2Loss.accumulate_gradients -- Gmain, Dmain with the distillation term, the lazy R1 phase -- exactly the calls
training_loop.py:321-331 makes) drive THIS repo's Generator and Discriminator modules through their public surface (`G.mapping(z=, c=, camera_angles=,
update_emas=)`, `G.synthesis(ws, camera_params, update_emas=, render_opts=, patch_params=)` returning a TensorGroup with `.img` / `.depth`,
`D(img, c, update_emas=, patch_params=, camera_angles=, predict_feat=)`), with the product's `conv2d_gradfix` / `upfirdn2d` bound where loss.py imports
them (:18-19).  Every parameter gradient each phase leaves behind must equal what the product's own loss (3dgp_b200/training/loss.py) leaves on the same
modules from the same RNG state -- which pins the product loss against the reference's, phase by phase.

With `learn_camera_dist=true` (the reference's default, configs/training/base.yaml:9) the reference phase also runs this repo's CameraAdaptor -- called with
the reference's TensorGroup, returning the product's, on which loss.py then does group arithmetic (`emd_regs + emd_regs.max() * 0.0`, member assignment) --
and its earth-mover / force-mean regularisers.  POT (`import ot`, loss.py:12) is absent from this image: oracle/pot_standin.py supplies `ot.dist` / `ot.emd2` as POT
documents them, solved exactly by scipy's assignment solver, so the product's sorted-matching restatement meets an independent solver inside the reference's own code.

CPU only (emulated C ABI).  Skipped where /root/reference is absent."""
import importlib
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

import abi_emulator as emu
from conftest import ROOT
from oracle import cases, ref_harness as rh
from util import l2rel

pytestmark = pytest.mark.skipif(not rh.available(), reason='the unmodified reference is only present in the build container')


@pytest.mark.parametrize('learn_camera_dist', [True], ids=['learn_camera_dist'])     # the reference default (configs/training/base.yaml:9); False is the same code minus the camera terms
def test_reference_training_phases_drive_our_modules_and_equal_our_loss(monkeypatch, learn_camera_dist):
    emu.install(monkeypatch)
    ns = rh.load()
    from oracle import pot_standin
    if 'ot' not in sys.modules:
        monkeypatch.setitem(sys.modules, 'ot', pot_standin)
    import src.training.loss as ref_loss
    monkeypatch.setattr(ref_loss, 'ot', pot_standin)
    cfgm = importlib.import_module('3dgp_b200.config')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    lossm = importlib.import_module('3dgp_b200.training.loss')
    monkeypatch.setattr(ref_loss, 'conv2d_gradfix', importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix'))
    monkeypatch.setattr(ref_loss, 'upfirdn2d', importlib.import_module('3dgp_b200.torch_utils.ops.upfirdn2d'))

    meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'networks_meta.json')))
    kw = dict(meta['net_kwargs']); kw.pop('learn_camera_dist', None)
    cfg = cfgm.make_config(**kw, kd_weight=1.0, batch_size=4, learn_camera_dist=learn_camera_dist)
    torch.manual_seed(0)
    G, D = cfgm.build_networks(cfg, 'cpu', fp32_D=True)
    # the camera adaptor keeps its constructor initialisation: the synthetic fill below makes its posterior almost constant (64 samples inside 3e-4), float32
    # ties between samples then leave the optimal transport plan non-unique and the two solvers pick different -- equally optimal -- sub-gradients
    ca_init = {k: v.clone() for k, v in G.synthesis.camera_adaptor.state_dict().items()} if learn_camera_dist else None
    G.load_state_dict(cases.fill_state_dict({k: tuple(v.shape) for k, v in G.state_dict().items()}, G.state_dict(), seed=100))
    D.load_state_dict(cases.fill_state_dict({k: tuple(v) for k, v in meta['D_keys'].items()}, D.state_dict(), seed=200))
    if learn_camera_dist:
        G.synthesis.camera_adaptor.load_state_dict(ca_init)
    G.train(); D.train()
    t = {k: torch.from_numpy(v) for k, v in cases.net_inputs(meta['net_kwargs']).items()}
    B, res = t['z'].shape[0], kw['img_resolution']
    g = torch.Generator().manual_seed(3)
    real_img, real_depth = torch.rand(B, 3, res, res, generator=g) * 2 - 1, torch.rand(B, 1, res, res, generator=g) * 2 - 1
    embs = torch.randn(B, kw['embedding_dim'], generator=g)

    L_ref = ref_loss.StyleGAN2Loss(cfg, 'cpu', G, D, r1_gamma=1.0)
    L_our = lossm.StyleGAN2Loss(cfg, 'cpu', G, D, r1_gamma=1.0)
    L_ref.progressive_update(5000); L_our.progressive_update(5000)
    assert L_ref.D_kd_weight == L_our.D_kd_weight > 0 and L_ref.patch_cfg.beta == L_our.patch_cfg.beta and L_ref.emd_multiplier == L_our.emd_multiplier == (0.5 if learn_camera_dist else 0.0)

    def data(d):
        real = d.EasyDict(img=real_img.clone(), depth=real_depth.clone(), c=t['c'].clone(), embs=embs.clone(), camera_angles=t['angles'].clone())
        gen = d.EasyDict(z=t['z'].clone(), c=t['c'].clone(), camera_angles_cond=None,      # camera_cond is off in the 3dgp configuration: the mapping ignores it
                         camera_params=d.TensorGroup(angles=t['angles'].clone(), fov=t['fov'].clone(), radius=t['radius'].clone(), look_at=t['look_at'].clone()))
        return real, gen

    phases = (('Gmain', G, 150_000), ('Dmain', D, 150_000), ('Dreg', D, 400_000))
    for phase, module, cur_nimg in phases:   # 150 kimg: blur sigma 2.5 (15 separable taps); 400 kimg: none
        got = {}
        for which, L, group in (('reference', L_ref, ns.dnnlib), ('product', L_our, dn)):
            G.requires_grad_(module is G); D.requires_grad_(module is D)
            for p in list(G.parameters()) + list(D.parameters()):
                p.grad = None
            torch.manual_seed(11); np.random.seed(11)
            G.synthesis.renderer.launch_counter = 0        # the ray-march draws its variates from a counter-based stream keyed by (seed, launch index)
            real, gen = data(group)
            L.accumulate_gradients(phase=phase, real_data=real, gen_data=gen, gain=(16 if phase == 'Dreg' else 1), cur_nimg=cur_nimg)
            got[which] = {n: p.grad.detach().clone() for n, p in module.named_parameters() if p.grad is not None}
        a, b = got['reference'], got['product']
        assert set(a) == set(b) and len(a) > 20, (phase, set(a) ^ set(b))
        worst = max((l2rel(b[n].numpy(), a[n].numpy()), n) for n in a if a[n].abs().max() > 0)
        assert worst[0] < 1e-5, (phase, worst)
        assert all(torch.equal(b[n], a[n]) for n in a if a[n].abs().max() == 0)


def test_tensor_group_has_the_reference_container_surface():
    """3dgp_b200/dnnlib.py::TensorGroup against src/dnnlib/util.py:66-175 on the operations loss.py / training_loop.py apply to camera-parameter groups:
    arithmetic with scalars and groups, whole-group reductions, slicing / split / repeat_interleave, nested groups."""
    ns = rh.load()
    dn = importlib.import_module('3dgp_b200.dnnlib')
    g = torch.Generator().manual_seed(0)
    angles, fov = torch.randn(5, 3, generator=g), torch.rand(5, generator=g) + 1
    mk = lambda T: T(angles=angles.clone(), fov=fov.clone(), nested=T(look_at=angles.clone() * 2))

    def flat(t):
        return [kv for k, v in t.items() for kv in (flat(v) if hasattr(v, 'items') else [(k, v)])]

    def same(x, y):
        if torch.is_tensor(x):
            assert torch.equal(x, y)
        elif isinstance(x, list) and x and hasattr(x[0], 'items'):
            assert len(x) == len(y)
            for i, j in zip(x, y):
                same(i, j)
        elif hasattr(x, 'items'):
            fx, fy = sorted(flat(x), key=lambda kv: kv[0]), sorted(flat(y), key=lambda kv: kv[0])      # the reference's `cat` walks a set of names: member order is not part of the contract
            assert [k for k, _ in fx] == [k for k, _ in fy]
            for (_, i), (_, j) in zip(fx, fy):
                same(i, j)
        else:
            assert x == y
    ops = [lambda t: t + 1.5, lambda t: 2 + t, lambda t: t - 0.5, lambda t: t * 3, lambda t: 0.0 * t, lambda t: t ** 2, lambda t: t + t, lambda t: t * t, lambda t: t - t,
           lambda t: t.max(), lambda t: t.sum(), lambda t: t.numel(), lambda t: t.reduce_mean(), lambda t: t.clone(), lambda t: t.float(), lambda t: t.to(torch.float64),
           lambda t: t.repeat_interleave(2, dim=0), lambda t: t.split(2), lambda t: t[1:3], lambda t: len(t), lambda t: t.shape, lambda t: t.keys(),
           lambda t: t.detach(), lambda t: t.cpu(), lambda t: t.clamp(-0.5, 0.5), lambda t: str(t.device)]
    flat_ops = [lambda t: t.reshape_each(lambda v: [v.shape[0], 1, -1]).permute(0, 2, 1), lambda t: t.mean(dim=0, keepdim=True), lambda t: t.reshape_each(lambda v: [v.shape[0], -1]),
                lambda t: [tuple(s) for s in t.shapes], lambda t: type(t).cat([t, t * 2], dim=0), lambda t: t.split(3)]
    a, b = mk(ns.dnnlib.TensorGroup), mk(dn.TensorGroup)
    for f in ops:
        same(f(a), f(b))
    fa, fb = (T(angles=angles.clone(), fov=fov.clone()) for T in (ns.dnnlib.TensorGroup, dn.TensorGroup))      # member-wise tensor methods: flat groups (the reference's do not recurse)
    for f in flat_ops:
        same(f(fa), f(fb))
    t = mk(dn.TensorGroup)
    t.fov = t.fov * 2.0                                  # member assignment, as loss.py:172-175 does on the rolled regulariser groups
    assert torch.equal(t.fov, fov * 2.0) and len(t) == 5
