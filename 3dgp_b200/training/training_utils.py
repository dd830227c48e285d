"""Patch sampling / schedules (reference src/training/training_utils.py)."""
import numpy as np
import torch
import torch.nn.functional as F


def linear_schedule(step, val_start, val_end, period, start_step=0):
    if step >= start_step + period:
        return val_end
    if step <= start_step:
        return val_start
    return val_start + (val_end - val_start) * (step - start_step) / period


def generate_coords(batch_size, img_size, device='cpu', align_corners=False):
    """[-1,1] grid, upper-left = (-1, 1), lower-right = (1, -1) (training_utils.py:147-168)."""
    if align_corners:
        row = torch.linspace(-1, 1, img_size, device=device).float()
    else:
        row = (torch.arange(0, img_size, device=device).float() / img_size) * 2 - 1
    x = row.view(1, -1).repeat(img_size, 1)
    y = -x.t()
    return torch.stack([x, y], dim=2).view(1, img_size, img_size, 2).repeat(batch_size, 1, 1, 1)


def compute_patch_coords(patch_params, resolution, align_corners=True, for_grid_sample=True):
    scales, offsets = patch_params['scales'], patch_params['offsets']
    B = scales.shape[0]
    coords = generate_coords(B, resolution, device=scales.device, align_corners=align_corners)
    coords = (coords + 1.0) * scales.view(B, 1, 1, 2) - 1.0 + offsets.view(B, 1, 1, 2) * 2.0
    if for_grid_sample:
        coords[:, :, :, 1] = -coords[:, :, :, 1]
    return coords


def extract_patches(x, patch_params, resolution):
    """Bilinear patch crop of the real images (training_utils.py:22-31)."""
    coords = compute_patch_coords(patch_params, resolution)
    return F.grid_sample(x, coords, mode='bilinear', align_corners=True)


def sample_patch_params(batch_size, patch_cfg, device='cpu'):
    """Beta / uniform / discrete-uniform patch scales shared inside minibatch-std groups (training_utils.py:57-143)."""
    g = patch_cfg.mbstd_group_size
    num_groups = batch_size // g
    if patch_cfg.distribution == 'beta':
        sx = np.random.beta(a=patch_cfg.alpha, b=patch_cfg.beta, size=num_groups) * (patch_cfg.max_scale - patch_cfg.min_scale) + patch_cfg.min_scale
    elif patch_cfg.distribution == 'uniform':
        sx = np.random.rand(num_groups) * (patch_cfg.max_scale - patch_cfg.min_scale) + patch_cfg.min_scale
    elif patch_cfg.distribution == 'discrete_uniform':        # configs/training/patch_discrete_uniform.yaml: the listed scales that lie inside the current range
        support = [v for v in patch_cfg.discrete_support if patch_cfg.min_scale <= v <= patch_cfg.max_scale]
        sx = np.random.choice(support, size=num_groups, replace=True).astype(np.float32)
    else:
        raise NotImplementedError(patch_cfg.distribution)
    sx = torch.from_numpy(sx).float().to(device)
    scales = torch.stack([sx, sx], dim=1)
    offsets = torch.rand(scales.shape, device=device) * (1.0 - scales)
    return {'scales': scales.repeat_interleave(g, dim=0), 'offsets': offsets.repeat_interleave(g, dim=0)}


def sample_random_c(batch_size, c_dim, device):
    """Uniform one-hot class labels (training_utils.py:207-214)."""
    c = torch.zeros(batch_size, c_dim, device=device)
    if c_dim > 0:
        c[torch.arange(batch_size, device=device), torch.randint(0, c_dim, (batch_size,), device=device)] = 1.0
    return c
