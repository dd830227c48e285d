"""Default experiment configuration = the reference's Hydra composition
configs/{camera/base+uniform, model/base+3dgp, training/base+patch_beta, dataset/base+imagenet}.yaml with the README overrides
(cmax=1024, cbase=65536), as nested EasyDicts.  The modules read exactly the keys listed in SURVEY.md 5 ("Config / flags"),
so an `experiment_config.yaml` produced by the reference's launcher loads unchanged (yaml.safe_load -> EasyDict.init_recursively)."""
import numpy as np

from .dnnlib import EasyDict


def make_config(cmax=1024, cbase=65536, tri_res=512, feat_dim=32, num_ray_steps=48, patch_res=64, img_resolution=256, c_dim=1000,
                use_depth=True, hid_dim=64, d_fmaps=1.0, w_dim=512, z_dim=512, depth_hid=64, embedding_dim=2048, kd_weight=1.0,
                batch_size=64, learn_camera_dist=False):
    camera = dict(
        ray=dict(start=0.75, end=1.25), fov=dict(dist='uniform', min=10.0, max=45.0),
        origin=dict(radius=dict(dist='normal', mean=1.0, std=0.0),
                    angles=dict(dist='uniform', yaw=dict(min=-1.57, max=1.57, mean=0.0, std=0.4),
                                pitch=dict(min=0.785398163, max=2.35619449, mean=1.57, std=0.2))),
        look_at=dict(radius=dict(dist='uniform', min=0.0, max=0.2),
                     angles=dict(dist='spherical_uniform', yaw=dict(min=-3.14159265, max=3.14159265), pitch=dict(min=0.0, max=3.14159265))),
        cube_scale=0.5, validate_viewing_frustum=False)
    patch = dict(enabled=True, patch_params_cond=True, min_scale_trg=patch_res / img_resolution, max_scale=1.0, anneal_kimg=10000,
                 resolution=patch_res, mbstd_group_size=4, distribution='beta', alpha=1.0, beta_val_start=0.001, beta_val_end=0.8)
    dataset = dict(c_dim=c_dim, resolution=img_resolution, white_back=False, last_back=False, embedding_dim=embedding_dim)
    generator = dict(
        fp32_only=True, cmax=cmax, cbase=cbase, fmaps=1.0, patch=patch, dataset=dataset, camera=camera, w_dim=w_dim, z_dim=z_dim, c_dim=c_dim,
        map_depth=2, use_inf_depth=True, has_view_cond=False, camera_cond=False, camera_cond_drop_p=0.0, camera_cond_spoof_p=0.5, density_bias=0.0,
        num_ray_steps=num_ray_steps, ray_marcher_type='classical', max_batch_res=128, use_full_box=False, architecture='skip', clamp_mode='softplus',
        nerf_noise_std_init=1.0, nerf_noise_kimg_growth=5000, use_noise=True,
        tri_plane=dict(res=tri_res, feat_dim=feat_dim, mlp=dict(n_layers=2, hid_dim=hid_dim)),
        depth_adaptor=dict(enabled=use_depth, kernel_size=5, hid_dim=depth_hid, num_hid_layers=3, out_strategy='random', selection_start_p=0.1,
                           anneal_kimg=10000, near_plane_offset_max_fraction=0.25, near_plane_offset_bias=-3.0, w_dim=w_dim, camera=camera),
        camera_adaptor=dict(enabled=learn_camera_dist, camera=camera, residual=False, lipschitz_weights=dict(enabled=False),
                            emd=dict(enabled=True, anneal_kimg=10000, num_samples=64, origin=2.0, radius=0.0, fov=0.0001, look_at=0.0001),
                            lr_multiplier=0.1, z_dim=z_dim, c_dim=c_dim, hid_dim=256, embed_dim=16,
                            adjust=dict(angles=True, radius=False, fov=True, look_at=True), force_mean_weight=10.0),
        optim=dict(kwargs=dict(lr=0.0025, betas=[0.0, 0.99], eps=1e-8, weight_decay=0.0)))
    discriminator = dict(fp32_only=False, c_dim=c_dim, cmax=cmax, cbase=cbase, fmaps=d_fmaps, patch=patch,
                         num_additional_start_blocks=int(np.log2(img_resolution // patch_res)), logits_clamp_val=1e7, mbstd_group_size=4,
                         camera_cond=False, camera_cond_drop_p=0.0, hyper_mod=True,
                         optim=dict(kwargs=dict(lr=0.002, betas=[0.0, 0.99], eps=1e-8, weight_decay=0.0)))
    loss_kwargs = dict(adv_loss_type='non_saturating', pl_weight=0.0, blur_init_sigma=10, blur_fade_kimg=200, gamma='auto',
                       kd=dict(discr=dict(weight=kd_weight, anneal_kimg=100000, loss_type='l2')))
    cfg = dict(camera=camera, dataset=dataset,
               model=dict(name='3dgp', generator=generator, discriminator=discriminator, loss_kwargs=loss_kwargs),
               training=dict(batch_size=batch_size, use_depth=use_depth, learn_camera_dist=learn_camera_dist, patch=patch, blur_real_depth_sigma=0.0))
    return EasyDict.init_recursively(cfg)


def set_reference_numerics():
    """training_loop.py:76-77: the reference trains with TF32 disabled for matmul and cuDNN convolutions."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def build_networks(cfg, device, fp32_D=False):
    """G, D as src/train.py:152-153,203,271-276 assembles them (G fp32-only, D with fp16 blocks unless fp32_D)."""
    set_reference_numerics()
    from .training.networks_epigraf import Generator
    from .training.networks_discriminator import Discriminator
    g = cfg.model.generator
    G = Generator(cfg=g, img_resolution=cfg.dataset.resolution, img_channels=3,
                  mapping_kwargs=dict(camera_cond=g.camera_cond, camera_cond_drop_p=g.camera_cond_drop_p, mean_camera_params=None),
                  fused_modconv_default='inference_only', num_fp16_res=0, conv_clamp=None)
    feat_dim = cfg.dataset.embedding_dim if cfg.model.loss_kwargs.kd.discr.weight > 0 else 0
    kw = dict(num_fp16_res=0, conv_clamp=None) if fp32_D else {}
    D = Discriminator(cfg=cfg.model.discriminator, input_resolution=cfg.training.patch.resolution, img_channels=3 + int(cfg.training.use_depth),
                      block_kwargs=dict(freeze_layers=0), mapping_kwargs={},
                      epilogue_kwargs=dict(mbstd_group_size=cfg.model.discriminator.mbstd_group_size, feat_predict_dim=feat_dim), **kw)
    return G.to(device), D.to(device)
