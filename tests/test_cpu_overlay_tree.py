"""INTEGRATION.md section B, executed literally: a COPY of the reference tree (`src/`) with this repo's files laid over it under the reference's own names --
`src/_lib.py`, `src/torch_utils/custom_ops.py`, `src/torch_utils/ops/*.py`, `src/training/{layers, networks_*, tri_plane_renderer}.py` -- and everything else
(`src/dnnlib`, `src/training/{loss, rendering_utils, training_utils}.py`, `src/torch_utils/{misc, persistence, training_stats}.py`, ...) left as the reference ships it.

Inside that tree, in a fresh interpreter: the networks are constructed the way src/train.py:157-203 + training_loop.py:113-114 construct them
(`dnnlib.util.construct_class_by_name(class_name='src.training.networks_epigraf.Generator', ...)`), loaded with the golden weights, run against the reference goldens, and
then driven by the tree's OWN `src/training/loss.py` (the reference file, which now imports this repo's `conv2d_gradfix` / `upfirdn2d` simply because they sit
where it looks for them).  The relative imports of this repo's modules must resolve against the reference's `dnnlib` package, `rendering_utils` and `training_utils`;
the names src/train.py imports from the overlaid files (`validate_image_plane`, train.py:29) must still exist.

CPU only (emulated C ABI, installed for the package name `src`); skipped where /root/reference is absent."""
import os
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT
from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason='the unmodified reference is only present in the build container')

OVERLAY = {
    '': ['_lib.py', 'build.py'],
    'torch_utils': ['custom_ops.py'],
    'torch_utils/ops': ['bias_act.py', 'upfirdn2d.py', 'filtered_lrelu.py', 'fma.py', 'conv2d_gradfix.py', 'conv2d_resample.py', 'tc.py', 'modconv.py', 'raymarch.py'],
    'training': ['layers.py', 'networks_epigraf.py', 'networks_stylegan2.py', 'networks_discriminator.py', 'networks_depth_adaptor.py', 'networks_camera_adaptor.py',
                 'tri_plane_renderer.py'],
}

SCRIPT = r'''
import importlib, json, os, sys, types
import numpy as np, torch, pytest
overlay, repo = sys.argv[1], sys.argv[2]
sys.path[:0] = [overlay, os.path.join(repo, 'tests'), repo]
from oracle import cases, ref_harness as rh, pot_standin
rh._install_stubs()                                   # omegaconf is imported by src/dnnlib/util.py for type hints only
sys.modules.setdefault('ot', pot_standin)             # POT: imported by src/training/loss.py
import abi_emulator as emu
mp = pytest.MonkeyPatch()
tc = emu.install(mp, package='src')
import src
assert os.path.dirname(src.__file__) == os.path.join(overlay, 'src'), src.__file__
from src import dnnlib
import src.training.loss as loss_mod, src.training.rendering_utils as ru, src.training.training_utils as tu
import src.training.networks_epigraf as ne, src.torch_utils.ops.conv2d_gradfix as cg
from src.training.tri_plane_renderer import validate_image_plane, ImportanceRenderer       # train.py:29
ref_lines = lambda m: open(m.__file__).read()
for mod_, rel in ((loss_mod, 'training/loss.py'), (ru, 'training/rendering_utils.py'), (tu, 'training/training_utils.py'), (dnnlib.util, 'dnnlib/util.py')):
    assert ref_lines(mod_) == open(os.path.join(rh.REF_ROOT, 'src', rel)).read(), rel           # the reference's own files, byte for byte
assert 'tcgen05' in ref_lines(cg) and hasattr(ne.SynthesisNetwork, 'forward') and 'forward_camera' in ref_lines(sys.modules['src.training.tri_plane_renderer'])
assert loss_mod.conv2d_gradfix is cg                   # bound by position in the tree, not by patching
assert validate_image_plane(fov=45.0, radius=1.0, scale=0.5, step=5e-2)
import warnings
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    import src.training.training_loop as tl            # the reference's loop module still imports with the overlaid ops in place (:25-35)
assert tl.conv2d_gradfix is cg and hasattr(cg, 'enabled') and tl.custom_ops.verbosity in ('none', 'brief', 'full') if hasattr(tl, 'custom_ops') else tl.conv2d_gradfix is cg

meta = json.load(open(os.path.join(repo, 'tests', 'golden', 'networks_meta.json')))
kw = meta['net_kwargs']
Gc, Dc, m = rh.make_cfg(**kw)
ED = dnnlib.EasyDict
G = dnnlib.util.construct_class_by_name(class_name='src.training.networks_epigraf.Generator', cfg=ED.init_recursively(Gc), img_resolution=m['img_resolution'], img_channels=3,
                                        mapping_kwargs=dict(camera_cond=False, camera_cond_drop_p=0.0, mean_camera_params=None),
                                        fused_modconv_default='inference_only', num_fp16_res=0, conv_clamp=None).train().requires_grad_(False)
D = dnnlib.util.construct_class_by_name(class_name='src.training.networks_discriminator.Discriminator', cfg=ED.init_recursively(Dc), input_resolution=m['patch_res'], img_channels=4,
                                        block_kwargs=dict(freeze_layers=0), mapping_kwargs={}, epilogue_kwargs=dict(mbstd_group_size=4, feat_predict_dim=m['embedding_dim']),
                                        num_fp16_res=0, conv_clamp=None).train().requires_grad_(False)
assert type(G).__name__ == 'Generator' and type(G).__mro__[1].__module__ == 'src.training.networks_epigraf' and isinstance(G.synthesis.renderer, ImportanceRenderer)   # [0] is persistence's wrapper class
assert {k: list(v.shape) for k, v in G.state_dict().items()} == meta['G_keys'] and {k: list(v.shape) for k, v in D.state_dict().items()} == meta['D_keys']
G.load_state_dict(cases.fill_state_dict({k: tuple(v) for k, v in meta['G_keys'].items()}, G.state_dict(), seed=100))
D.load_state_dict(cases.fill_state_dict({k: tuple(v) for k, v in meta['D_keys'].items()}, D.state_dict(), seed=200))

gold = np.load(os.path.join(repo, 'tests', 'golden', 'networks.npz'))
t = {k: torch.from_numpy(v) for k, v in cases.net_inputs(kw).items()}
B = t['z'].shape[0]
cam = dnnlib.TensorGroup(angles=t['angles'], fov=t['fov'], radius=t['radius'], look_at=t['look_at'])
pp = dict(scales=t['patch_scales'], offsets=t['patch_offsets'])
maxrel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
G.synthesis.nerf_noise_std = 0.0
with torch.no_grad():
    ws = G.mapping(t['z'], t['c'])
    o = G.synthesis(ws, cam, patch_params=pp, noise_mode='random', layer_noises=[torch.from_numpy(n) for n in cases.layer_noises(kw, B)],
                    render_opts=dict(concat_depth=True, return_depth=True, u_coarse=t['u_coarse'], u_fine=t['u_fine'], depth_head_idx=torch.from_numpy(cases.depth_heads(B)), mlp_mode=0))
    assert isinstance(o, dnnlib.TensorGroup)           # the reference's container class, returned by this repo's module
    e = (maxrel(ws.numpy(), gold['G/ws']), maxrel(o.img.numpy(), gold['G/train/img']), maxrel(o.depth.numpy(), gold['G/train/depth']))
    assert max(e) < 1e-4, e
    lg, ft = D(torch.from_numpy(gold['G/train/img']), t['c'], patch_params=pp, camera_angles=t['angles'], predict_feat=True)
    assert maxrel(lg.numpy(), gold['D/logits']) < 1e-4 and maxrel(ft.numpy(), gold['D/feats']) < 1e-4

# the reference's inference helpers and its metric loop's call pattern (src/training/inference_utils.py, metric_utils.py:303-319, 343-344) on this Generator
import src.training.inference_utils as iu
G.eval()
with torch.no_grad():
    z2, c2 = t['z'][:2], t['c'][:2]
    cp = iu.sample_posterior_camera_params(G, z2, c2)
    img = G(z=z2[:1], c=c2[:1], camera_params=cp[:1], camera_angles_cond=cp.angles[:1], noise_mode='const', render_opts=dict(cut_quantile=0.0))
    assert torch.is_tensor(img) and tuple(img.shape) == (1, 3, kw['img_resolution'], kw['img_resolution']) and torch.isfinite(img).all()
    frames = iu.generate(ED(batch_size=1), G, ws=G.mapping(z=z2, c=c2)[:1], camera_params=cp[:1], verbose=False, render_opts=dict(return_depth=True, return_depth_adapted=True))
    assert isinstance(frames, dnnlib.TensorGroup) and set(frames.keys()) == {'img', 'depth', 'depth_adapted'} and len(frames) == 1
    assert tuple(frames.depth.shape) == (1, 1, kw['img_resolution'], kw['img_resolution']) and float(frames.img.min()) >= 0.0 and float(frames.img.max()) <= 1.0
    mean_cam = iu.approximate_mean_camera_params(G, num_samples=16)
    assert tuple(mean_cam.angles.shape) == (1, 3)
G.train()

# the tree's own loss.py (reference file) runs its phases on these modules
cfgm = importlib.import_module('3dgp_b200.config')      # only for the composed experiment configuration (a plain nested dict)
cfg = ED.init_recursively(json.loads(json.dumps(cfgm.make_config(**{k: v for k, v in kw.items() if k != 'learn_camera_dist'}, kd_weight=1.0, batch_size=4))))
loss = dnnlib.util.construct_class_by_name(class_name='src.training.loss.StyleGAN2Loss', device='cpu', G=G, D=D, augment_pipe=None, cfg=cfg, r1_gamma=1.0)   # training_loop.py:186
res = kw['img_resolution']
g = torch.Generator().manual_seed(1)
for phase, module in (('Dmain', D),):          # Dmain runs the generator too (run_G under no_grad); Gmain / R1 on these modules: tests/test_cpu_reference_loss.py
    G.requires_grad_(module is G); D.requires_grad_(module is D)
    for p in module.parameters():
        p.grad = None
    real = ED(img=torch.rand(B, 3, res, res, generator=g) * 2 - 1, depth=torch.rand(B, 1, res, res, generator=g) * 2 - 1, c=t['c'].clone(),
              embs=torch.randn(B, m['embedding_dim'], generator=g), camera_angles=t['angles'].clone())
    gen = ED(z=t['z'].clone(), c=t['c'].clone(), camera_angles_cond=None, camera_params=dnnlib.TensorGroup(angles=t['angles'].clone(), fov=t['fov'].clone(), radius=t['radius'].clone(), look_at=t['look_at'].clone()))
    loss.accumulate_gradients(phase=phase, real_data=real, gen_data=gen, gain=1, cur_nimg=400000)
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    assert len(grads) > 20 and all(torch.isfinite(x).all() for x in grads) and sum(float(x.abs().sum()) for x in grads) > 0, phase
# training_loop.py:478-484: snapshot = CPU copies of the networks, pickled through the reference's persistence hooks (the overlaid classes carry the decorator)
import copy, pickle
from src.torch_utils import persistence
assert persistence.is_persistent(G) and persistence.is_persistent(G.synthesis.tri_plane_decoder.b8.conv1) and not persistence.is_persistent(D)   # as in the reference
snap = dict(G=copy.deepcopy(G).eval().requires_grad_(False).cpu(), G_ema=copy.deepcopy(G).eval().requires_grad_(False).cpu(), training_set_kwargs=dict(path='x.zip'))
with open(sys.argv[3], 'wb') as f:
    pickle.dump(snap, f)
torch.save({k: v.clone() for k, v in G.state_dict().items()}, sys.argv[3] + '.sd')
print('OVERLAY OK', e)
'''


LOADER = r'''
import os, pickle, sys
import torch
overlay, repo, path = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path[:0] = [overlay, os.path.join(repo, 'tests'), repo]
from oracle import ref_harness as rh
rh._install_stubs()
import src.torch_utils.persistence                      # all that pickle needs to find `_reconstruct_persistent_obj`; no network module has been imported
assert not any(m.startswith('src.training.networks') for m in sys.modules)
with open(path, 'rb') as f:
    snap = pickle.load(f)
G = snap['G_ema']
assert type(G).__name__ == 'Generator' and type(G).__mro__[1].__module__.startswith('_imported_module_'), [c.__module__ for c in type(G).__mro__[:3]]   # rebuilt from the embedded source
sd = torch.load(path + '.sd')
got = G.state_dict()
assert list(got) == list(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
assert G.init_kwargs['img_resolution'] == G.img_resolution and not G.training
# the same file through this repo's reader (3dgp_b200/legacy.py: no exec of the embedded source)
import importlib
lg = importlib.import_module('3dgp_b200.legacy')
back = lg.load_network_pkl(path, names=('G_ema',))
assert all(torch.equal(back['G_ema'].state_dict()[k], sd[k]) for k in sd)
print('SNAPSHOT OK', len(sd))
'''


def _make_overlay(tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_overlay', os.path.join(ROOT, 'tools', 'make_overlay.py'))
    tool = importlib.util.module_from_spec(spec); spec.loader.exec_module(tool)
    out, written = tool.install(rh.REF_ROOT, dest=str(tmp_path / 'overlay'))
    assert out == str(tmp_path / 'overlay' / 'src') and len(written) == sum(len(v) for v in OVERLAY.values())
    assert sorted(written) == sorted((f'{k}/{n}' if k else n) for k, v in OVERLAY.items() for n in v)
    for rel in written:                                    # no relative import survives (persistence.py execs module sources in anonymous modules)
        text = open(os.path.join(out, rel)).read()
        assert not [l for l in text.splitlines() if l.lstrip().startswith(('from .', 'import .'))], rel
    assert os.path.exists(os.path.join(out, 'csrc', 'conv_tc.cu')) and os.path.exists(str(tmp_path / 'overlay' / 'include' / 'gp3d_b200.h'))
    return str(tmp_path / 'overlay')


def test_reference_tree_with_our_files_laid_over_it(tmp_path):
    overlay = _make_overlay(tmp_path)                      # tools/make_overlay.py: the overlay recipe of INTEGRATION.md as a command
    script = tmp_path / 'drive.py'
    script.write_text(SCRIPT)
    r = subprocess.run([sys.executable, str(script), overlay, ROOT, str(tmp_path / 'snapshot.pkl')], capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0 and 'OVERLAY OK' in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
    # a snapshot written by the reference's persistence machinery in that process loads in a FRESH interpreter that has not imported the network modules:
    # persistence.py re-creates the classes from the module source stored in the pickle (the reason the overlaid files use absolute imports)
    loader = tmp_path / 'load.py'
    loader.write_text(LOADER)
    r = subprocess.run([sys.executable, str(loader), overlay, ROOT, str(tmp_path / 'snapshot.pkl')], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0 and 'SNAPSHOT OK' in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
