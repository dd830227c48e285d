"""Depth adaptor (reference src/training/networks_depth_adaptor.py:21-99): 3 x (5x5 conv, lrelu) with a shared 1x1 head
after every layer; one of the 4 candidate depth maps is picked per sample."""
import numpy as np
import torch

from .layers import Conv2dLayer
from .training_utils import linear_schedule


class DepthAdaptor(torch.nn.Module):
    def __init__(self, cfg, min_depth, max_depth):
        super().__init__()
        self.cfg, self.min_depth, self.max_depth = cfg, min_depth, max_depth
        self.depth_range = max_depth - min_depth
        dims = [1] + [cfg.hid_dim] * cfg.num_hid_layers
        self.layers = torch.nn.ModuleList([Conv2dLayer(i, o, cfg.kernel_size, activation='lrelu') for i, o in zip(dims[:-1], dims[1:])])
        self.head = Conv2dLayer(dims[-1], 1, 1, activation='linear') if len(self.layers) > 0 else None
        self.register_buffer('progress_coef', torch.tensor([0.0]))
        self.near_plane_offset_raw = torch.nn.Parameter(torch.tensor([cfg.near_plane_offset_bias]).float())

    def get_near_plane_offset(self, w):
        return self.near_plane_offset_raw.repeat(len(w)).sigmoid() * self.cfg.near_plane_offset_max_fraction * self.depth_range

    def normalize(self, x, w):
        near = (self.min_depth + self.get_near_plane_offset(w)).view(len(x), 1, 1, 1)
        return (x - 0.5 * (self.max_depth + near)) / ((self.max_depth - near) + 1e-12) * 2.0

    def progressive_update(self, cur_kimg):
        self.progress_coef.data = torch.tensor(linear_schedule(cur_kimg, 0.0, 1.0, self.cfg.anneal_kimg)).to(self.progress_coef.device)

    @property
    def start_p(self):
        return (1.0 / (self.cfg.num_hid_layers + 1) * (1 - self.progress_coef) + self.cfg.selection_start_p * self.progress_coef).item()

    def forward(self, depth_map, w, head_idx=None):
        """head_idx: optional [B] int64 overriding the np.random.choice draw of :91 (parity runs)."""
        x = self.normalize(depth_map, w)
        outs = [x]
        for layer in self.layers:
            x = layer(x)
            outs.append(self.head(x))
        outs = torch.stack(outs).transpose(0, 1)
        B, n = outs.shape[:2]
        if self.cfg.out_strategy == 'last':
            return outs[:, -1] + 0.0 * outs.max()
        if self.cfg.out_strategy == 'mean':
            return outs.mean(dim=1)
        if self.cfg.out_strategy == 'random':
            if head_idx is None:
                if self.training:
                    idx = np.arange(n)
                    slope = (1 - n * self.start_p) * 2 / (n * (n - 1))
                    head_idx = torch.from_numpy(np.random.choice(idx, size=(B,), p=idx * slope + self.start_p))
                else:       # eval: the last head for every sample (networks_depth_adaptor.py:93-94); plain slicing, capturable into a CUDA graph
                    return outs[:, n - 1] + 0.0 * outs.max()
            return outs[torch.arange(B, device=outs.device), head_idx.to(outs.device)] + 0.0 * outs.max()
        raise NotImplementedError(self.cfg.out_strategy)
