"""Camera helpers (reference src/training/rendering_utils.py:72-218, 270-285)."""
import numpy as np
import torch

from ..dnnlib import TensorGroup


def normalize(x, dim=-1):
    return x / torch.norm(x, dim=dim, keepdim=True)


def spherical2cartesian(rotation, pitch, radius=1.0):
    """Camera convention of the reference (:270-285): x = r sin(pitch) sin(-yaw), y = r cos(pitch), z = r sin(pitch) cos(yaw)."""
    x = radius * torch.sin(pitch) * torch.sin(-rotation)
    y = radius * torch.cos(pitch)
    z = radius * torch.sin(pitch) * torch.cos(rotation)
    return torch.stack([x, y, z], dim=-1)


def compute_cam2world_matrix(camera_params):
    """angles [B,3] (yaw,pitch,roll), radius [B], look_at [B,3] (yaw,pitch,radius) -> cam2world [B,4,4]  (:194-218)."""
    origins = spherical2cartesian(camera_params.angles[:, 0], camera_params.angles[:, 1], camera_params.radius)
    look_at = spherical2cartesian(camera_params.look_at[:, 0], camera_params.look_at[:, 1], camera_params.look_at[:, 2])
    fwd = normalize(normalize(look_at - origins))
    B = fwd.shape[0]
    up = torch.zeros_like(fwd)          # (0, 1, 0) built on the device: no host -> device copy, so the call can be captured into a CUDA graph
    up[:, 1] = 1.0
    left = normalize(torch.cross(up, fwd, dim=-1))
    up = normalize(torch.cross(fwd, left, dim=-1))
    rot = torch.eye(4, device=fwd.device).unsqueeze(0).repeat(B, 1, 1)
    rot[:, :3, :3] = torch.stack((-left, up, -fwd), dim=-1)
    tr = torch.eye(4, device=fwd.device).unsqueeze(0).repeat(B, 1, 1)
    tr[:, :3, 3] = origins
    return tr @ rot


def _angles(cfg, B, device):
    if cfg.dist == 'uniform':
        yaw = torch.rand((B, 1), device=device) * (cfg.yaw.max - cfg.yaw.min) + cfg.yaw.min
        pitch = torch.rand((B, 1), device=device) * (cfg.pitch.max - cfg.pitch.min) + cfg.pitch.min
    elif cfg.dist == 'normal':
        yaw = torch.randn((B, 1), device=device) * cfg.yaw.std + cfg.yaw.mean
        pitch = torch.randn((B, 1), device=device) * cfg.pitch.std + cfg.pitch.mean
    elif cfg.dist == 'spherical_uniform':
        yr, yc = cfg.yaw.max - cfg.yaw.min, 0.5 * (cfg.yaw.max + cfg.yaw.min)
        pr, pc = cfg.pitch.max - cfg.pitch.min, 0.5 * (cfg.pitch.max + cfg.pitch.min)
        yaw = (torch.rand((B, 1), device=device) - 0.5) * yr + yc
        v = (torch.rand((B, 1), device=device) - 0.5) * pr + pc
        pitch = torch.arccos(1 - 2 * torch.clamp(v / np.pi, 1e-5, 1 - 1e-5))
    elif cfg.dist == 'truncnorm':       # configs/camera/base.yaml: centred on the middle of the allowed range, clipped to it
        yaw = _truncnorm((cfg.yaw.max + cfg.yaw.min) * 0.5, cfg.yaw.std, cfg.yaw.min, cfg.yaw.max, B, device).unsqueeze(1)
        pitch = _truncnorm((cfg.pitch.max + cfg.pitch.min) * 0.5, cfg.pitch.std, cfg.pitch.min, cfg.pitch.max, B, device).unsqueeze(1)
    else:
        raise NotImplementedError(f'camera angle distribution `{cfg.dist}` (built: uniform / normal / truncnorm / spherical_uniform)')
    pitch = torch.clamp(pitch, 1e-5, np.pi - 1e-5)
    return torch.cat([yaw, pitch, torch.zeros_like(yaw)], dim=1)


def _truncnorm(mean, std, lo, hi, B, device):
    """Normal(mean, std) restricted to [lo, hi], drawn as the reference draws it (rendering_utils.py:140-146: scipy's sampler on numpy's global stream)."""
    from scipy.stats import truncnorm
    x = truncnorm.rvs(a=(lo - mean) / std, b=(hi - mean) / std, loc=mean, scale=std, size=(B,))
    return torch.from_numpy(x).float().to(device)


def _scalar(cfg, B, device):
    if cfg.dist == 'normal':
        assert cfg.std == 0.0, 'Scalar must be bounded'
        return torch.full([B], float(cfg.mean), device=device)
    if cfg.dist == 'truncnorm':
        return _truncnorm(cfg.mean, cfg.std, cfg.min, cfg.max, B, device)
    if cfg.dist == 'uniform':
        return torch.rand(B, device=device) * (cfg.max - cfg.min) + cfg.min
    raise NotImplementedError(cfg.dist)


def sample_camera_params(cfg, batch_size, device='cpu', origin_angles=None):
    """Prior camera sampling (:150-156) for the uniform / normal / truncnorm / spherical_uniform distributions of configs/camera (same draws from the same RNG state as the reference)."""
    angles = _angles(cfg.origin.angles, batch_size, device) if origin_angles is None else origin_angles
    fov = _scalar(cfg.fov, batch_size, device)
    radius = _scalar(cfg.origin.radius, batch_size, device)
    la = _angles(cfg.look_at.angles, batch_size, device)
    look_at = torch.cat([la[:, [0, 1]], _scalar(cfg.look_at.radius, batch_size, device).unsqueeze(1)], dim=1)
    return TensorGroup(angles=angles, fov=fov, radius=radius, look_at=look_at)


def get_mean_sampling_value(cfg):
    """Mean of a scalar prior (rendering_utils.py:170-176): fov, origin radius."""
    if cfg.dist in ('normal', 'truncnorm'):
        return cfg.mean
    if cfg.dist == 'uniform':
        return (cfg.max + cfg.min) / 2
    raise NotImplementedError(f'mean of distribution `{cfg.dist}`')


def get_mean_angles_values(angles_cfg):
    """Mean (yaw, pitch, roll) of an origin-angle prior (rendering_utils.py:180-190)."""
    if angles_cfg.dist == 'normal':
        return [angles_cfg.yaw.mean, angles_cfg.pitch.mean, 0.0]
    if angles_cfg.dist in ('spherical_uniform', 'truncnorm', 'uniform'):
        return [(angles_cfg.yaw.max + angles_cfg.yaw.min) * 0.5, (angles_cfg.pitch.max + angles_cfg.pitch.min) * 0.5, 0.0]
    raise NotImplementedError(f'mean of camera angle distribution `{angles_cfg.dist}`')
