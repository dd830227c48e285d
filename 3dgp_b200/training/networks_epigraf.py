"""3DGP / EpiGRAF generator (reference src/training/networks_epigraf.py): mapping -> StyleGAN2 tri-plane decoder ->
fused volumetric render -> depth adaptor.  Same classes, constructor arguments, attributes and state-dict keys.

The tri-plane image is kept in channels-last storage ([B, P, P, 96]) from the last ToRGB/upsample on, which is the
channel-minor layout the fused ray-march kernel gathers from (one 128-byte line per bilinear tap).
"""
import math

import numpy as np
import torch

from ..dnnlib import EasyDict, TensorGroup
from .layers import FullyConnectedLayer, MappingNetwork
from .networks_depth_adaptor import DepthAdaptor
from .networks_camera_adaptor import CameraAdaptor
from .networks_stylegan2 import SynthesisBlock
from .rendering_utils import compute_cam2world_matrix
from .training_utils import linear_schedule
from .tri_plane_renderer import ImportanceRenderer, sample_rays


class TriPlaneMLP(torch.nn.Module):
    """feat_dim -> hid -> (out_dim + 1); evaluated inside the fused kernel, this module only owns the parameters
    (state-dict keys `model.{0,1}.{weight,bias}`, networks_epigraf.py:29-68)."""

    def __init__(self, cfg, out_dim):
        super().__init__()
        self.cfg, self.out_dim = cfg, out_dim
        if cfg.tri_plane.mlp.n_layers != 2 or cfg.has_view_cond:
            raise NotImplementedError('fused ray-march is built for the 2-layer, view-independent tri-plane MLP of configs/model/3dgp.yaml')
        self.backbone_out_dim = 1 + out_dim
        self.dims = [cfg.tri_plane.feat_dim, cfg.tri_plane.mlp.hid_dim, self.backbone_out_dim]
        self.model = torch.nn.Sequential(FullyConnectedLayer(self.dims[0], self.dims[1], activation='lrelu'),
                                         FullyConnectedLayer(self.dims[1], self.dims[2], activation='linear'))

    def forward(self, x):
        """[B, 3, M, feat_dim] -> {'rgb': [B,M,out_dim], 'sigma': [B,M,1]} -- unfused path for callers that sample the
        field directly (compute_densities); the renderer does not call this."""
        B, _, M, C = x.shape
        x = self.model(x.mean(dim=1).reshape(B * M, C)).view(B, M, self.backbone_out_dim)
        return {'rgb': x[..., :-1], 'sigma': x[:, :, [-1]]}


class SynthesisBlocksSequence(torch.nn.Module):
    def __init__(self, cfg, in_resolution, out_resolution, in_channels, out_channels, num_fp16_res=4, **block_kwargs):
        assert in_resolution == 0 or (in_resolution >= 4 and math.log2(in_resolution).is_integer())
        assert out_resolution >= 4 and math.log2(out_resolution).is_integer() and in_resolution < out_resolution
        super().__init__()
        self.cfg, self.out_resolution, self.in_channels, self.out_channels, self.num_fp16_res = cfg, out_resolution, in_channels, out_channels, num_fp16_res
        lo = 2 if in_resolution == 0 else int(np.log2(in_resolution)) + 1
        hi = int(np.log2(out_resolution))
        self.block_resolutions = [2 ** i for i in range(lo, hi + 1)]
        ch = {res: min(int(cfg.cbase * cfg.fmaps) // res, cfg.cmax) for res in self.block_resolutions}
        fp16_res = max(2 ** (hi + 1 - num_fp16_res), 8)
        self.num_ws = 0
        for i, res in enumerate(self.block_resolutions):
            cin = ch[res // 2] if i > 0 else in_channels
            is_last = (res == out_resolution)
            block = SynthesisBlock(cin, ch[res], w_dim=cfg.w_dim, resolution=res, img_channels=out_channels, is_last=is_last,
                                   use_fp16=(res >= fp16_res), **block_kwargs)
            self.num_ws += block.num_conv
            if is_last:
                self.num_ws += block.num_torgb
            setattr(self, f'b{res}', block)

    def forward(self, ws, x=None, layer_noises=None, **block_kwargs):
        ws = ws.to(torch.float32)
        block_ws, w_idx = [], 0
        for res in self.block_resolutions:
            block = getattr(self, f'b{res}')
            block_ws.append(ws.narrow(1, w_idx, block.num_conv + block.num_torgb))
            w_idx += block.num_conv
        img = None
        nz = list(layer_noises) if layer_noises is not None else None
        for res, cur in zip(self.block_resolutions, block_ws):
            block = getattr(self, f'b{res}')
            ln = None
            if nz is not None:
                ln, nz = nz[:block.num_conv], nz[block.num_conv:]
            x, img = block(x, img, cur, layer_noises=ln, **block_kwargs)
        return img


class SynthesisNetwork(torch.nn.Module):
    def __init__(self, cfg, img_resolution, img_channels, **synthesis_seq_kwargs):
        super().__init__()
        self.cfg, self.img_resolution, self.img_channels = cfg, img_resolution, img_channels
        self.tri_plane_decoder = SynthesisBlocksSequence(cfg=cfg, in_resolution=0, out_resolution=cfg.tri_plane.res, in_channels=0,
                                                         out_channels=cfg.tri_plane.feat_dim * 3, architecture='skip', use_noise=cfg.use_noise,
                                                         **synthesis_seq_kwargs)
        self.tri_plane_mlp = TriPlaneMLP(cfg, out_dim=img_channels)
        self.num_ws = self.tri_plane_decoder.num_ws
        self.nerf_noise_std = 0.0
        self.train_resolution = cfg.patch.resolution if cfg.patch.enabled else img_resolution
        self.test_resolution = img_resolution
        self.renderer = ImportanceRenderer(ray_marcher_type=cfg.ray_marcher_type)
        self.depth_adaptor = DepthAdaptor(cfg.depth_adaptor, min_depth=cfg.camera.ray.start, max_depth=cfg.camera.ray.end) if cfg.depth_adaptor.enabled else None
        self.camera_adaptor = CameraAdaptor(cfg.camera_adaptor) if cfg.camera_adaptor.enabled else None        # networks_epigraf.py:174-177
        self._default_render_options = EasyDict(max_batch_res=cfg.max_batch_res, return_depth=False, return_depth_adapted=False,
                                                return_weights=False, concat_depth=False, cut_quantile=0.0, density_bias=cfg.density_bias)

    def progressive_update(self, cur_kimg):
        self.nerf_noise_std = linear_schedule(cur_kimg, self.cfg.nerf_noise_std_init, 0.0, self.cfg.nerf_noise_kimg_growth)
        if self.depth_adaptor is not None:
            self.depth_adaptor.progressive_update(cur_kimg)

    @torch.no_grad()
    def compute_densities(self, ws, coords, max_batch_res=32, **block_kwargs):
        """sigma of the radiance field at world points `coords` [B, M, 3] -> [B, M, 1] (networks_epigraf.py:196-208; scripts/extract_geometry.py samples a
        volume with it).  Not a rendering: no rays, no compositing -- each point reads its three bilinear plane taps (points outside the cube read zeros),
        the taps are averaged and go through the tri-plane MLP; evaluated in chunks of max_batch_res^3 points."""
        F_, P = self.cfg.tri_plane.feat_dim, self.cfg.tri_plane.res
        dec = self.tri_plane_decoder(ws[:, :self.tri_plane_decoder.num_ws], **block_kwargs)
        B = dec.shape[0]
        planes = dec[:, :3 * F_].reshape(B * 3, F_, P, P)                   # plane-major channel groups: xy, xz, yz
        out = []
        for pts in coords.split(max_batch_res ** 3, dim=1):
            q = pts / self.cfg.camera.cube_scale
            uv = torch.stack([q[..., [0, 1]], q[..., [0, 2]], q[..., [1, 2]]], dim=1).reshape(B * 3, 1, -1, 2)
            taps = torch.nn.functional.grid_sample(planes, uv, mode='bilinear', align_corners=True)       # [B*3, F, 1, m]
            feats = taps.reshape(B, 3, F_, -1).permute(0, 1, 3, 2)                                         # [B, 3, m, F]
            out.append(self.tri_plane_mlp(feats)['sigma'])
        return torch.cat(out, dim=1)

    def forward(self, ws, camera_params, patch_params=None, render_opts={}, **block_kwargs):
        """ws [B,num_ws,w_dim]; camera_params {angles [B,3], fov [B], radius [B], look_at [B,3]}; patch_params {scales, offsets}.
        render_opts may carry `u_coarse`, `u_fine`, `sn_coarse`, `sn_fine`, `depth_head_idx` (parity runs)."""
        ro = EasyDict(**{**self._default_render_options, **render_opts})
        cfg = self.cfg
        B, N = ws.shape[0], cfg.num_ray_steps
        dec = self.tri_plane_decoder(ws[:, :self.tri_plane_decoder.num_ws], **block_kwargs)       # [B, 3*feat, P, P]
        F_, P = cfg.tri_plane.feat_dim, cfg.tri_plane.res
        dec = dec[:, :3 * F_].contiguous(memory_format=torch.channels_last)
        planes = dec.view(B, 3, F_, P, P)
        h = w = (self.train_resolution if self.training else self.test_resolution)
        noise_std = self.nerf_noise_std if self.training else 0.0
        c2w = compute_cam2world_matrix(camera_params)
        if cfg.use_full_box:
            raise NotImplementedError('use_full_box=true (ray/box intersection bounds) is not on the 3dgp path')
        opts = EasyDict(box_size=cfg.camera.cube_scale * 2, num_proposal_steps=N, clamp_mode=cfg.get('clamp_mode', 'softplus'),
                        use_inf_depth=cfg.use_inf_depth, ray_start=cfg.camera.ray.start, ray_end=cfg.camera.ray.end, num_fine_steps=N,
                        density_noise=noise_std, last_back=cfg.dataset.last_back, white_back=cfg.dataset.white_back,
                        max_batch_res=ro.max_batch_res, cut_quantile=ro.cut_quantile, density_bias=ro.density_bias)
        for k in ('u_coarse', 'u_fine', 'sn_coarse', 'sn_fine', 'mlp_mode', 'seed'):
            if k in ro:
                opts[k] = ro[k]
        # no run_batchwise chunking (networks_epigraf.py:232-240): the fused kernel never materialises per-sample tensors
        if opts.get('mlp_mode', 2) == 0:      # first-generation fp32 SIMT kernels take explicit rays
            ray_o, ray_d = sample_rays(c2w, fov=camera_params.fov, resolution=(h, w), patch_params=patch_params, device=ws.device)
            feats, depths, _w, _t = self.renderer(planes, self.tri_plane_mlp, ray_o, ray_d, opts)
        else:                                 # ray generation (tri_plane_renderer.py:487-527) fused into the render launch
            feats, depths, _w, _t = self.renderer.forward_camera(planes, self.tri_plane_mlp, c2w, camera_params.fov, (h, w), patch_params, opts)
        img = feats.reshape(B, h, w, self.img_channels).permute(0, 3, 1, 2).contiguous()
        depth = depths.reshape(B, 1, h, w)
        depth_adapted = None
        if self.depth_adaptor is not None:
            depth_adapted = self.depth_adaptor(depth, ws[:, 0], head_idx=ro.get('depth_head_idx'))
            img = torch.cat([img, depth_adapted], dim=1) if ro.concat_depth else img + 0.0 * depth_adapted.max()
        if ro.return_depth or ro.return_depth_adapted:
            out = TensorGroup(img=img)
            if ro.return_depth:
                out.depth = depth
            if ro.return_depth_adapted:
                out.depth_adapted = depth_adapted
            return out
        return img


class Generator(torch.nn.Module):
    def __init__(self, cfg, img_resolution, img_channels, mapping_kwargs={}, **synthesis_kwargs):
        super().__init__()
        self.cfg = cfg
        self.z_dim, self.c_dim, self.w_dim = cfg.z_dim, cfg.c_dim, cfg.w_dim
        self.img_resolution, self.img_channels = img_resolution, img_channels
        self.synthesis = SynthesisNetwork(cfg=cfg, img_resolution=img_resolution, img_channels=img_channels, **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=cfg.z_dim, c_dim=cfg.c_dim, w_dim=cfg.w_dim, num_ws=self.num_ws, camera_raw_scalars=True,
                                      num_layers=cfg.map_depth, **mapping_kwargs)

    def progressive_update(self, cur_kimg):
        self.synthesis.progressive_update(cur_kimg)

    def forward(self, z, c, camera_params, camera_angles_cond=None, truncation_psi=1, truncation_cutoff=None, update_emas=False, **synthesis_kwargs):
        ws = self.mapping(z, c, camera_angles=camera_angles_cond, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.synthesis(ws, camera_params=camera_params, update_emas=update_emas, **synthesis_kwargs)
