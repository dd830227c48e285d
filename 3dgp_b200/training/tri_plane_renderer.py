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
