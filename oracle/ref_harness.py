"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference (snap-research/3dgp at
/root/reference) inside the build container so that

  * oracle/restated.py (our CPU restatement) can be pinned against it, and
  * tests/golden/*.npz fixtures can be generated (oracle/make_golden.py).

/root/reference does not exist on the GPU box: nothing under tests/ -m gpu, bench.py or
__graft_entry__.smoke() imports this module.  It is never imported by the product package.

The reference needs `omegaconf` only for type hints and one isinstance() check
(src/dnnlib/util.py:35,58-59), so a ~10 line in-memory stub is installed.
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get('GP3D_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, 'src', 'training'))


def _install_stubs():
    if 'omegaconf' not in sys.modules:
        m = types.ModuleType('omegaconf')

        class DictConfig(dict):
            pass

        class OmegaConf:
            @staticmethod
            def to_yaml(cfg):
                return repr(cfg)

            @staticmethod
            def set_struct(cfg, flag):
                return None

        m.DictConfig = DictConfig
        m.OmegaConf = OmegaConf
        sys.modules['omegaconf'] = m


_loaded = {}


def load():
    """Returns a namespace with the reference modules needed for the hot path."""
    if _loaded:
        return _loaded['ns']
    assert available(), f'reference not found at {REF_ROOT}'
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from src import dnnlib
        from src.torch_utils.ops import bias_act, upfirdn2d, filtered_lrelu, conv2d_resample, conv2d_gradfix, fma
        from src.training import networks_epigraf, networks_discriminator, networks_stylegan2, layers
        from src.training import tri_plane_renderer, rendering_utils, training_utils
    ns = types.SimpleNamespace(
        dnnlib=dnnlib, bias_act=bias_act, upfirdn2d=upfirdn2d, filtered_lrelu=filtered_lrelu,
        conv2d_resample=conv2d_resample, conv2d_gradfix=conv2d_gradfix, fma=fma,
        networks_epigraf=networks_epigraf, networks_discriminator=networks_discriminator,
        networks_stylegan2=networks_stylegan2, layers=layers, tri_plane_renderer=tri_plane_renderer,
        rendering_utils=rendering_utils, training_utils=training_utils)
    _loaded['ns'] = ns
    return ns


# ----------------------------------------------------------------------------------------------
# Config transcription of configs/{camera/base+uniform, model/base+3dgp|epigraf,
# training/base+patch_beta, dataset/base+imagenet}.yaml (reference file:line in SURVEY.md 5).

def make_cfg(cmax=1024, cbase=65536, tri_res=512, feat_dim=32, num_ray_steps=48, patch_res=64,
             img_resolution=256, c_dim=1000, use_depth=True, learn_camera_dist=False,
             hid_dim=64, d_fmaps=1.0, w_dim=512, z_dim=512, depth_hid=64, embedding_dim=2048):
    """Returns (G_cfg, D_cfg, meta) as plain nested dicts (converted to EasyDict by callers)."""
    camera = dict(
        ray=dict(start=0.75, end=1.25),
        fov=dict(dist='uniform', min=10.0, max=45.0),
        origin=dict(radius=dict(dist='normal', mean=1.0, std=0.0),
                    angles=dict(dist='uniform', yaw=dict(min=-1.57, max=1.57, mean=0.0, std=0.4),
                                pitch=dict(min=0.785398163, max=2.35619449, mean=1.57, std=0.2))),
        look_at=dict(radius=dict(dist='uniform', min=0.0, max=0.2),
                     angles=dict(dist='spherical_uniform', yaw=dict(min=-3.14159265, max=3.14159265),
                                 pitch=dict(min=0.0, max=3.14159265))),
        cube_scale=0.5, validate_viewing_frustum=False)
    patch = dict(enabled=True, patch_params_cond=True, min_scale_trg=patch_res / img_resolution, max_scale=1.0,
                 anneal_kimg=10000, resolution=patch_res, mbstd_group_size=4, distribution='beta', alpha=1.0,
                 beta_val_start=0.001, beta_val_end=0.8)
    dataset = dict(c_dim=c_dim, resolution=img_resolution, white_back=False, last_back=False,
                   embedding_dim=embedding_dim)
    G = dict(
        fp32_only=True, cmax=cmax, cbase=cbase, fmaps=1.0, patch=patch, dataset=dataset, camera=camera,
        w_dim=w_dim, z_dim=z_dim, c_dim=c_dim, map_depth=2, use_inf_depth=True, has_view_cond=False,
        camera_cond=False, camera_cond_drop_p=0.0, camera_cond_spoof_p=0.5, density_bias=0.0,
        num_ray_steps=num_ray_steps, ray_marcher_type='classical', max_batch_res=128, use_full_box=False,
        architecture='skip', clamp_mode='softplus', nerf_noise_std_init=1.0, nerf_noise_kimg_growth=5000,
        use_noise=True,
        tri_plane=dict(res=tri_res, feat_dim=feat_dim, mlp=dict(n_layers=2, hid_dim=hid_dim)),
        depth_adaptor=dict(enabled=use_depth, kernel_size=5, hid_dim=depth_hid, num_hid_layers=3,
                           out_strategy='random', selection_start_p=0.1, anneal_kimg=10000,
                           near_plane_offset_max_fraction=0.25, near_plane_offset_bias=-3.0, w_dim=w_dim,
                           camera=camera),
        camera_adaptor=dict(enabled=learn_camera_dist, camera=camera, residual=False,
                            lipschitz_weights=dict(enabled=False),
                            emd=dict(enabled=True, anneal_kimg=10000, num_samples=64, origin=2.0, radius=0.0,
                                     fov=0.0001, look_at=0.0001),
                            lr_multiplier=0.1, z_dim=z_dim, c_dim=c_dim, hid_dim=256, embed_dim=16,
                            adjust=dict(angles=True, radius=False, fov=True, look_at=True),
                            force_mean_weight=10.0))
    num_add = int(np.log2(img_resolution // patch_res))
    D = dict(fp32_only=False, c_dim=c_dim, cmax=cmax, cbase=cbase, fmaps=d_fmaps, patch=patch,
             num_additional_start_blocks=num_add, logits_clamp_val=1e7, mbstd_group_size=4, camera_cond=False,
             camera_cond_drop_p=0.0, hyper_mod=True)
    meta = dict(img_resolution=img_resolution, patch_res=patch_res, embedding_dim=embedding_dim,
                use_depth=use_depth)
    return G, D, meta


def build_reference_G(G_cfg, img_resolution, seed=0):
    ns = load()
    cfg = ns.dnnlib.EasyDict.init_recursively(G_cfg)
    torch.manual_seed(seed)
    # train.py:203,271-276 -> class + fp32_only kwargs
    G = ns.networks_epigraf.Generator(cfg=cfg, img_resolution=img_resolution, img_channels=3,
                                      mapping_kwargs=dict(camera_cond=False, camera_cond_drop_p=0.0,
                                                          mean_camera_params=None),
                                      fused_modconv_default='inference_only', num_fp16_res=0, conv_clamp=None)
    return G


def build_reference_D(D_cfg, patch_res, use_depth=True, embedding_dim=0, seed=1, fp32=True):
    ns = load()
    cfg = ns.dnnlib.EasyDict.init_recursively(D_cfg)
    torch.manual_seed(seed)
    kw = dict(num_fp16_res=0, conv_clamp=None) if fp32 else {}
    D = ns.networks_discriminator.Discriminator(
        cfg=cfg, input_resolution=patch_res, img_channels=3 + int(use_depth),
        block_kwargs=dict(freeze_layers=0), mapping_kwargs={},
        epilogue_kwargs=dict(mbstd_group_size=4, feat_predict_dim=embedding_dim), **kw)
    return D


@contextlib.contextmanager
def injected_rng(rand_like=(), rand=(), randn=(), randn_like=()):
    """Replaces torch.rand_like / torch.rand / torch.randn / torch.randn_like by queues of pre-generated
    tensors so that the reference's stochastic renderer consumes *given* noise (SURVEY.md 8a 'RNG')."""
    queues = dict(rand_like=list(rand_like), rand=list(rand), randn=list(randn), randn_like=list(randn_like))
    orig = dict(rand_like=torch.rand_like, rand=torch.rand, randn=torch.randn, randn_like=torch.randn_like)

    def make(name):
        def f(*a, **k):
            q = queues[name]
            assert len(q) > 0, f'injected_rng: {name} queue exhausted'
            t = q.pop(0)
            return t.clone()
        return f

    try:
        for name in orig:
            if len(queues[name]) > 0:
                setattr(torch, name, make(name))
        yield queues
    finally:
        for name, fn in orig.items():
            setattr(torch, name, fn)
