"""Ray generation + the volumetric renderer front-end (reference src/training/tri_plane_renderer.py).

`ImportanceRenderer.forward(planes, decoder, ray_origins, ray_directions, rendering_options)` keeps the reference's
signature and 4-tuple result (:126-170) but executes as ONE fused sm_100a kernel (csrc/raymarch_fwd.cu) -- coarse pass,
importance sampling, fine pass, depth merge and compositing -- instead of ~25 torch ops over per-sample tensors.
"""
from typing import Dict, Tuple, Union

import numpy as np
import torch

from ..torch_utils.ops import raymarch
from .rendering_utils import normalize


def sample_rays(c2w, fov, resolution: Tuple[int, int], patch_params: Dict = None, device=None):
    """World-space ray origins / directions of a (patch of a) pinhole camera, tri_plane_renderer.py:487-527."""
    B = len(c2w)
    cb = 1 if (patch_params is None and type(fov) is float) else B
    w, h = resolution
    x, y = torch.meshgrid(torch.linspace(-1, 1, w, device=device), torch.linspace(1, -1, h, device=device), indexing='ij')
    x = x.T.flatten().unsqueeze(0).repeat(cb, 1)
    y = y.T.flatten().unsqueeze(0).repeat(cb, 1)
    if patch_params is not None:
        ps, po = patch_params['scales'], patch_params['offsets']
        x = (x + 1.0) * ps[:, 0].view(B, 1) - 1.0 + po[:, 0].view(B, 1) * 2.0
        y = (y + 1.0) * ps[:, 1].view(B, 1) - 1.0 + po[:, 1].view(B, 1) * 2.0
    fov = fov if isinstance(fov, torch.Tensor) else torch.tensor([fov], device=device)
    fov_rad = fov.unsqueeze(1).expand(cb, 1) / 360 * 2 * np.pi
    z = -torch.ones((cb, h * w), device=device) / torch.tan(fov_rad * 0.5)
    d_cam = normalize(torch.stack([x, y, z], dim=2), dim=2)
    if cb == 1:
        d_cam = d_cam.repeat(B, 1, 1)
    d_world = torch.bmm(c2w[..., :3, :3], d_cam.reshape(B, -1, 3).permute(0, 2, 1)).permute(0, 2, 1).reshape(B, h * w, 3)
    ho = torch.zeros((B, 4, h * w), device=device)
    ho[:, 3, :] = 1
    o_world = torch.bmm(c2w, ho).permute(0, 2, 1).reshape(B, h * w, 4)[..., :3]
    return o_world, d_world


def get_ray_limits_box(rays_o, rays_d, box_size):
    """Entry / exit distance of each ray through the axis-aligned cube [-box_size/2, box_size/2]^3 (slab test), (-1, -2) where the ray misses it;
    rays [..., 3] -> two [..., 1] tensors.  The launcher's geometry checks use it (tri_plane_renderer.py:409-461); the render kernel itself marches the
    fixed [ray_start, ray_end] interval of configs/model/3dgp.yaml."""
    o = rays_o.detach().reshape(-1, 3)
    inv = 1 / rays_d.detach().reshape(-1, 3)
    half = box_size / 2
    t_a, t_b = (-half - o) * inv, (half - o) * inv                # per-axis crossings of the two faces
    near, far = torch.minimum(t_a, t_b), torch.maximum(t_a, t_b)
    # slabs are intersected in x, y, z order; a ray is rejected when an interval starts after the running one has ended (or the reverse) -- comparisons
    # with NaN (a ray inside a face plane and parallel to it) are false, exactly like the sequential form
    t0, t1 = near[:, 0], far[:, 0]
    hit = torch.ones_like(t0, dtype=torch.bool)
    for ax in (1, 2):
        hit &= ~((t0 > far[:, ax]) | (near[:, ax] > t1))
        t0, t1 = torch.max(t0, near[:, ax]), torch.min(t1, far[:, ax])
    t0 = torch.where(hit, t0, torch.full_like(t0, -1))
    t1 = torch.where(hit, t1, torch.full_like(t1, -2))
    shape = tuple(rays_o.shape[:-1]) + (1,)
    return t0.reshape(shape), t1.reshape(shape)


def validate_image_plane(fov, radius, scale=1.0, step=1e-2, device='cpu'):
    """True when, from every camera position on the sphere of `radius` (a yaw x pitch grid of pi/2/step angles each), the four corner rays of the image
    plane pass through the scene cube of half-extent `scale` -- the configuration check of src/train.py:211-215 (tri_plane_renderer.py:531-556)."""
    from ..dnnlib import TensorGroup
    from .rendering_utils import compute_cam2world_matrix
    n = int((np.pi / 2) / step)
    yaw, pitch = torch.meshgrid(torch.linspace(0, np.pi * 2, steps=n, device=device), torch.linspace(0, np.pi, steps=n, device=device), indexing='ij')
    angles = torch.stack([yaw.reshape(-1), pitch.clamp(1e-7, np.pi - 1e-7).reshape(-1), torch.zeros(n * n, device=device)], dim=1)
    cams = TensorGroup(angles=angles, radius=torch.full([n * n], float(radius), device=device), fov=torch.full([n * n], float(fov), device=device),
                       look_at=torch.zeros_like(angles))
    ray_o, ray_d = sample_rays(compute_cam2world_matrix(cams), fov=cams.fov, resolution=(2, 2), patch_params=None, device=device)
    t_in, t_out = get_ray_limits_box(ray_o, ray_d, box_size=scale * 2)
    return bool((t_out > t_in).all().item())


class ImportanceRenderer(torch.nn.Module):
    def __init__(self, ray_marcher_type: str):
        super().__init__()
        if ray_marcher_type != 'classical':
            raise NotImplementedError('only ray_marcher_type=classical (configs/model/{3dgp,epigraf}.yaml) is built')
        self.ray_marcher_type = ray_marcher_type
        self.launch_counter = 0   # Philox offset: a fresh stream per call

    def forward(self, planes, decoder, ray_origins, ray_directions, rendering_options):
        """planes [B,3,C,P,P]; decoder: TriPlaneMLP (2 FullyConnectedLayers); rays [B,R,3].
        rendering_options may carry injected variates `u_coarse`, `u_fine`, `sn_coarse`, `sn_fine` ([B,R,N]) for parity runs."""
        ro = rendering_options
        if ro.get('cut_quantile', 0.0) > 0.0:
            raise NotImplementedError('cut_quantile > 0 is a visualisation-only option and is not built')
        if ro['num_fine_steps'] != ro['num_proposal_steps']:
            raise NotImplementedError('the fused kernel assumes num_fine_steps == num_proposal_steps (networks_epigraf.py:228-229)')
        fc0, fc1 = decoder.model[0], decoder.model[1]
        self.launch_counter += 1
        rgb, depth, wsum, tfin = raymarch.render_rays(
            planes, fc0.weight, fc0.bias, fc1.weight, fc1.bias, ray_origins, ray_directions,
            num_steps=ro['num_proposal_steps'], ray_start=ro['ray_start'], ray_end=ro['ray_end'], box_size=ro['box_size'],
            u_coarse=ro.get('u_coarse'), u_fine=ro.get('u_fine'), sn_coarse=ro.get('sn_coarse'), sn_fine=ro.get('sn_fine'),
            density_noise=ro.get('density_noise', 0.0), use_inf_depth=ro.get('use_inf_depth', True), last_back=ro.get('last_back', False),
            white_back_end_idx=ro.get('white_back_end_idx', 0), clamp_mode=ro.get('clamp_mode', 'softplus'),
            mlp_mode=ro.get('mlp_mode', 2), seed=ro.get('seed', 0), offset=self.launch_counter)
        return rgb, depth, wsum, tfin

    def forward_camera(self, planes, decoder, c2w, fov, resolution, patch_params, rendering_options):
        """sample_rays (:487-527) + forward (:126-170) as ONE launch: the rays of the pinhole cameras `c2w` [B,4,4] / `fov` [B] (degrees) over a
        `resolution` = (h, w) grid (optionally the patch `patch_params` = {scales, offsets} of it) are generated inside the kernel; CTAs own 4 x 4
        pixel tiles.  Same 4-tuple result; differentiable w.r.t. planes, decoder parameters, c2w and fov."""
        ro = rendering_options
        if ro.get('cut_quantile', 0.0) > 0.0:
            raise NotImplementedError('cut_quantile > 0 is a visualisation-only option and is not built')
        if ro['num_fine_steps'] != ro['num_proposal_steps']:
            raise NotImplementedError('the fused kernel assumes num_fine_steps == num_proposal_steps (networks_epigraf.py:228-229)')
        fc0, fc1 = decoder.model[0], decoder.model[1]
        self.launch_counter += 1
        B = planes.shape[0]
        fov_t = fov if isinstance(fov, torch.Tensor) else torch.full([B], float(fov), device=planes.device)
        ps = po = None
        if patch_params is not None and len(patch_params) > 0:
            ps, po = patch_params['scales'], patch_params['offsets']
        return raymarch.render_camera(
            planes, fc0.weight, fc0.bias, fc1.weight, fc1.bias, c2w, fov_t, resolution, ps, po,
            num_steps=ro['num_proposal_steps'], ray_start=ro['ray_start'], ray_end=ro['ray_end'], box_size=ro['box_size'],
            u_coarse=ro.get('u_coarse'), u_fine=ro.get('u_fine'), sn_coarse=ro.get('sn_coarse'), sn_fine=ro.get('sn_fine'),
            density_noise=ro.get('density_noise', 0.0), use_inf_depth=ro.get('use_inf_depth', True), last_back=ro.get('last_back', False),
            white_back_end_idx=ro.get('white_back_end_idx', 0), clamp_mode=ro.get('clamp_mode', 'softplus'),
            mlp_mode=ro.get('mlp_mode', 2) or 2, seed=ro.get('seed', 0), offset=self.launch_counter)
