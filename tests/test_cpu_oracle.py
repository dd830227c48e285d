"""CPU suite: the oracle restatement (oracle/restated.py) against the committed golden vectors that were produced by
running the unmodified reference (oracle/make_golden.py).  Also re-pins against the live reference when
/root/reference is present (build container only)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cases, restated as R, ref_harness
from util import maxrel

TOL = 2e-6   # fp32 re-association noise (observed <= 2e-6 at generation time, tests/golden/oracle_pin_report.json)


@pytest.mark.parametrize('name,kw', cases.upfirdn2d_cases(), ids=[c[0] for c in cases.upfirdn2d_cases()])
def test_upfirdn2d_oracle_vs_golden(golden, name, kw):
    x, f = cases.upfirdn2d_inputs(name, kw)
    y = R.upfirdn2d(x, f, up=kw['up'], down=kw['down'], padding=kw['padding'], flip_filter=kw['flip_filter'], gain=kw['gain'])
    g = golden('upfirdn2d')[name]
    assert y.shape == g.shape
    if kw.get('integer', False):
        assert np.array_equal(y, g)          # bit-exact indexing (and exact arithmetic on dyadic data)
    else:
        assert maxrel(y, g) < TOL


def test_upfirdn2d_out_size_formula():
    # integer formula of upfirdn2d.cpp:35-36
    for (n, up, down, p0, p1, fs) in [(16, 2, 1, 2, 1, 4), (17, 1, 1, 1, 1, 4), (16, 1, 2, 1, 1, 4), (7, 3, 2, 2, 3, 5), (1, 2, 1, 2, 1, 4)]:
        assert R.upfirdn2d_out_size(n, up, down, p0, p1, fs) == (n * up + p0 + p1 - fs + down) // down


@pytest.mark.parametrize('name,kw', cases.bias_act_cases(), ids=[c[0] for c in cases.bias_act_cases()])
def test_bias_act_oracle_vs_golden(golden, name, kw):
    x, b = cases.bias_act_inputs(name, kw)
    y = R.bias_act(x, b, dim=kw['dim'], act=kw['act'], alpha=kw.get('alpha'), gain=kw.get('gain'), clamp=kw.get('clamp'))
    assert maxrel(y, golden('bias_act')[name + '/y']) < TOL


@pytest.mark.parametrize('name,kw', cases.filtered_lrelu_cases(), ids=[c[0] for c in cases.filtered_lrelu_cases()])
def test_filtered_lrelu_oracle_vs_golden(golden, name, kw):
    x, fu, fd, b = cases.filtered_lrelu_inputs(name, kw)
    y = R.filtered_lrelu(x, fu, fd, b, up=kw['up'], down=kw['down'], padding=kw['padding'], gain=kw['gain'], slope=kw['slope'], clamp=kw['clamp'])
    assert maxrel(y, golden('filtered_lrelu')[name + '/y']) < TOL


@pytest.mark.parametrize('name,kw', cases.render_cases(), ids=[c[0] for c in cases.render_cases()])
def test_render_oracle_vs_golden(golden, name, kw):
    inp = cases.render_inputs(name, kw)
    t = {k: torch.from_numpy(v) for k, v in inp.items()}
    rgb, depth, wsum, T = R.render(t['planes'], t['w1'], t['b1'], t['w2'], t['b2'], t['ray_o'], t['ray_d'], t['u_coarse'], t['u_fine'],
                                   kw['ray_start'], kw['ray_end'], kw['box_half'], kw['N'], sn_coarse=t.get('sn_coarse'), sn_fine=t.get('sn_fine'),
                                   noise_std=kw.get('noise_std', 0.0), use_inf_depth=kw.get('use_inf_depth', True), last_back=kw.get('last_back', False),
                                   white_back_end_idx=kw.get('white_back_end_idx', 0), clamp_mode=kw.get('clamp_mode', 'softplus'))
    g = golden('render')
    assert maxrel(rgb, g[name + '/rgb']) < TOL
    assert maxrel(depth, g[name + '/depth']) < TOL
    assert maxrel(wsum, g[name + '/wsum']) < TOL
    assert maxrel(T, g[name + '/tfinal']) < TOL


@pytest.mark.parametrize('variant', ['small', 'wide'])
def test_generator_discriminator_oracle_vs_golden(golden, variant):
    from conftest import ROOT
    tag = 'networks' if variant == 'small' else 'networks_wide'
    meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', tag + '_meta.json')))
    kw = meta['net_kwargs']
    pr = (lambda a: np.asarray(a)) if variant == 'small' else (lambda a: cases.grad_probe(np.asarray(a)))
    Gc, Dc, m = ref_harness.make_cfg(**kw)
    sdG = cases.fill_state_dict({k: tuple(v) for k, v in meta['G_keys'].items()}, _structural_buffers(meta['G_keys']), seed=100)
    sdD = cases.fill_state_dict({k: tuple(v) for k, v in meta['D_keys'].items()}, _structural_buffers(meta['D_keys']), seed=200)
    inp = cases.net_inputs(kw)
    t = {k: torch.from_numpy(v) for k, v in inp.items()}
    B = t['z'].shape[0]
    g = golden(tag)
    num_ws = g['G/ws'].shape[1]
    ws = R.mapping_network(sdG, 'mapping.', t['z'], t['c'], num_ws)
    assert maxrel(ws, g['G/ws']) < TOL
    noises = [torch.from_numpy(n) for n in cases.layer_noises(kw, B)]
    o = R.generator_synthesis(sdG, Gc, ws, t['angles'], t['fov'], t['radius'], t['look_at'], kw['patch_res'], t['patch_scales'], t['patch_offsets'],
                              t['u_coarse'], t['u_fine'], noise_mode='random', noises=noises, fused_modconv=False,
                              depth_head_idx=torch.from_numpy(cases.depth_heads(B)))
    assert maxrel(o['planes'].flatten()[::31], g['G/train/planes_probe']) < 5e-6
    assert maxrel(o['img'], g['G/train/img']) < 5e-6
    assert maxrel(o['depth'], g['G/train/depth']) < 5e-6
    block_res = [2 ** i for i in range(int(np.log2(kw['img_resolution'])), 2, -1)]
    lg, f = R.discriminator(sdD, torch.from_numpy(g['G/train/img']), t['c'], t['patch_scales'], t['patch_offsets'], block_res,
                            Dc['num_additional_start_blocks'], predict_feat=True)
    assert maxrel(lg, g['D/logits']) < 1e-5
    assert maxrel(f, g['D/feats']) < 1e-5


def _structural_buffers(keys):
    """resample_filter / fourier_coefs / progress_coef buffers as the reference constructors create them
    (upfirdn2d.setup_filter([1,3,3,1]) -- networks_stylegan2.py:116; construct_log_spaced_freqs -- layers.py:339-350)."""
    out = {}
    for k, shp in keys.items():
        if k.endswith('resample_filter'):
            out[k] = torch.from_numpy(R.setup_filter([1, 3, 3, 1]))
        elif k.endswith('fourier_coefs'):
            n = shp[0]
            out[k] = (2.0 ** torch.arange(n).float() / (2 ** n)) * np.pi
        elif k.endswith('progress_coef'):
            out[k] = torch.zeros(1)
    return out


@pytest.mark.skipif(not ref_harness.available(), reason='reference checkout not present (GPU box)')
def test_oracle_repinned_against_live_reference():
    """Build container only: run one op + the renderer through the imported reference and compare the restatement."""
    ns = ref_harness.load()
    name, kw = cases.upfirdn2d_cases()[1]
    x, f = cases.upfirdn2d_inputs(name, kw)
    y_ref = ns.upfirdn2d._upfirdn2d_ref(torch.from_numpy(x), torch.from_numpy(f), up=kw['up'], down=kw['down'], padding=kw['padding'],
                                        flip_filter=kw['flip_filter'], gain=kw['gain']).numpy()
    assert maxrel(R.upfirdn2d(x, f, up=kw['up'], down=kw['down'], padding=kw['padding'], flip_filter=kw['flip_filter'], gain=kw['gain']), y_ref) < TOL
    rep = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'oracle_pin_report.json')))
    assert max(rep.values()) < 2e-4 and len(rep) > 60


def test_fast_fir_path_used_for_cpu_timing_matches_the_restatement():
    name, kw = cases.upfirdn2d_cases()[5]
    x, f = cases.upfirdn2d_inputs(name, kw)
    a = R.upfirdn2d(x, f, up=kw['up'], down=kw['down'], padding=kw['padding'], flip_filter=kw['flip_filter'], gain=kw['gain'])
    b = R._upfirdn2d_aten(x, f, kw['up'], kw['down'], kw['padding'], kw['flip_filter'], kw['gain'])
    assert a.shape == b.shape and maxrel(a, b) < TOL
    name, kw = cases.upfirdn2d_cases()[8]           # separable 61-tap blur
    x, f = cases.upfirdn2d_inputs(name, kw)
    a = R.upfirdn2d(x, f, up=kw['up'], down=kw['down'], padding=kw['padding'], flip_filter=kw['flip_filter'], gain=kw['gain'])
    b = R._upfirdn2d_aten(x, f, kw['up'], kw['down'], kw['padding'], kw['flip_filter'], kw['gain'])
    assert a.shape == b.shape and maxrel(a, b) < TOL


@pytest.mark.parametrize('variant', ['small', 'wide'])
def test_differentiable_oracle_gradients_vs_golden(golden, variant):
    """oracle/restated.py with DIFFERENTIABLE=True (the arithmetic bench.py's executed CPU training step runs, oracle/train_step.py) against the
    reference's own autograd: Gmain loss and parameter gradients through G and D, D loss incl. the R1 double backward."""
    from conftest import ROOT
    tag = 'networks' if variant == 'small' else 'networks_wide'
    meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', tag + '_meta.json')))
    kw = meta['net_kwargs']
    Gc, Dc, m = ref_harness.make_cfg(**kw)
    sdG = cases.fill_state_dict({k: tuple(v) for k, v in meta['G_keys'].items()}, _structural_buffers(meta['G_keys']), seed=100)
    sdD = cases.fill_state_dict({k: tuple(v) for k, v in meta['D_keys'].items()}, _structural_buffers(meta['D_keys']), seed=200)
    t = {k: torch.from_numpy(v) for k, v in cases.net_inputs(kw).items()}
    B = t['z'].shape[0]
    g = golden(tag)
    pr = (lambda a: a) if variant == 'small' else cases.grad_probe
    old = (R.DIFFERENTIABLE, R.FAST_FIR)
    R.DIFFERENTIABLE, R.FAST_FIR = True, True
    try:
        namesG, namesD = cases.probe_params('G', variant), cases.probe_params('D', variant)
        for n in namesG:
            sdG[n].requires_grad_(True)
        for n in namesD:
            sdD[n].requires_grad_(True)
        noises = [torch.from_numpy(n) for n in cases.layer_noises(kw, B)]
        ws = R.mapping_network(sdG, 'mapping.', t['z'], t['c'], g['G/ws'].shape[1])
        o = R.generator_synthesis(sdG, Gc, ws, t['angles'], t['fov'], t['radius'], t['look_at'], kw['patch_res'], t['patch_scales'], t['patch_offsets'],
                                  t['u_coarse'], t['u_fine'], noise_mode='random', noises=noises, fused_modconv=False,
                                  depth_head_idx=torch.from_numpy(cases.depth_heads(B)))
        top = kw['img_resolution']
        block_res = [2 ** i for i in range(int(np.log2(top)), 2, -1)]
        lg, _ = R.discriminator(sdD, o['img'], t['c'], t['patch_scales'], t['patch_offsets'], block_res, Dc['num_additional_start_blocks'])
        lossG = torch.nn.functional.softplus(-lg).mean()
        assert abs(lossG.item() - float(g['G/loss'][0])) < 1e-5 * max(1.0, abs(float(g['G/loss'][0])))
        gs = torch.autograd.grad(lossG, [sdG[n] for n in namesG])
        for n, gr in zip(namesG, gs):
            assert maxrel(pr(gr.numpy()), g['G/grad/' + n]) < 2e-4, n
        img = torch.from_numpy(g['G/train/img']).requires_grad_(True)
        lg, _ = R.discriminator(sdD, img, t['c'], t['patch_scales'], t['patch_offsets'], block_res, Dc['num_additional_start_blocks'], predict_feat=True)
        r1 = torch.autograd.grad([lg.sum()], [img], create_graph=True)[0]
        assert maxrel(r1.detach().numpy(), g['D/r1_grads']) < 1e-4
        loss = torch.nn.functional.softplus(-lg).mean() + r1.square().sum([1, 2, 3]).mean() * 0.5
        gs = torch.autograd.grad(loss, [sdD[n] for n in namesD])
        for n, gr in zip(namesD, gs):
            assert maxrel(pr(gr.numpy()), g['D/grad/' + n]) < 2e-4, n
    finally:
        R.DIFFERENTIABLE, R.FAST_FIR = old
