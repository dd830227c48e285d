"""One data-parallel optimisation step of the reference training loop (src/training/training_loop.py:291-366):
phases Gmain / Dmain (+ lazy Dreg), gradient accumulation, ONE flattened all-reduce per phase (the only collective on
the hot path, :335-344) with the `/world + nan_to_num(+-1e5)` epilogue fused into one kernel, Adam, G_ema lerp.

One process per GPU; `torch.distributed` (NCCL over NVLink on the B200 box, gloo in CPU tests) carries the all-reduce.
"""
import copy

import torch
import torch.distributed as dist

from .. import _lib
from ..dnnlib import EasyDict


def allreduce_gradients(params, world_size, group=None):
    """flat = cat(grads); all_reduce(SUM); flat = nan_to_num(flat / world, 0, 1e5, -1e5); scatter back (training_loop.py:335-344).
    On CUDA the divide + nan_to_num runs as one pass (gp3d_grad_epilogue) instead of two torch ops."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return 0
    flat = torch.cat([p.grad.flatten() for p in params])
    if world_size > 1:
        dist.all_reduce(flat, group=group)
    if flat.is_cuda:
        with torch.cuda.device(flat.device):
            rc = _lib.lib().gp3d_grad_epilogue(flat.data_ptr(), flat.numel(), 1.0 / world_size, 1e5, -1e5, _lib.stream_ptr())
        _lib.check(rc, 'grad_epilogue')
    else:   # host-side logic under test with gloo (no GPU): same arithmetic, torch ops
        flat = torch.nan_to_num(flat / world_size, nan=0, posinf=1e5, neginf=-1e5)
    for p, g in zip(params, flat.split([p.numel() for p in params])):
        p.grad = g.reshape(p.shape)
    return flat.numel()


class Trainer:
    """Holds G, D, G_ema, the loss and both Adam optimisers with the reference's lazy-regularisation scaling
    (training_loop.py:190-205: lr *= mb_ratio, betas ** mb_ratio with mb_ratio = interval / (interval + 1))."""

    def __init__(self, G, D, loss, cfg, rank=0, world_size=1, D_reg_interval=16, ema_kimg=10.0, ema_rampup=0.05, batch_size=None, micro_batch=None):
        self.G, self.D, self.loss, self.cfg, self.rank, self.world_size = G, D, loss, cfg, rank, world_size
        if world_size > 1:   # training_loop.py:176-179: every rank starts from rank 0's parameters and buffers
            for module in (G, D):
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t.data, src=0)
        self.G_ema = copy.deepcopy(G).eval().requires_grad_(False)
        gk, dk = dict(cfg.model.generator.optim.kwargs), dict(cfg.model.discriminator.optim.kwargs)
        self.G_opt = torch.optim.Adam(G.parameters(), **gk)
        mb = D_reg_interval / (D_reg_interval + 1) if D_reg_interval else 1.0
        dk['lr'] = dk['lr'] * mb
        dk['betas'] = [b ** mb for b in dk['betas']]
        self.D_opt = torch.optim.Adam(D.parameters(), **dk)
        self.D_reg_interval = D_reg_interval
        self.ema_kimg, self.ema_rampup, self.batch_size = ema_kimg, ema_rampup, batch_size
        self.micro_batch = micro_batch
        self.cur_nimg = 0
        self.it = 0

    def _phase(self, name, module, opt, real, gen, gain, render_opts=None):
        opt.zero_grad(set_to_none=True)
        module.requires_grad_(True)
        stats = {}
        for r_mb, g_mb in self._micro_batches(real, gen):     # gradient accumulation, training_loop.py:329-330
            stats = self.loss.accumulate_gradients(phase=name, real_data=r_mb, gen_data=g_mb, gain=gain, cur_nimg=self.cur_nimg, render_opts=render_opts)
        module.requires_grad_(False)
        allreduce_gradients([p for p in module.parameters() if p.numel() > 0], self.world_size)
        opt.step()
        return stats

    def _micro_batches(self, real, gen):
        n = len(gen.z)
        mb = self.micro_batch or n
        if mb >= n:
            yield real, gen
            return
        sl = lambda d, a, b: EasyDict(**{k: (v[a:b] if hasattr(v, '__getitem__') else v) for k, v in d.items()})
        for a in range(0, n, mb):
            yield sl(real, a, a + mb), sl(gen, a, a + mb)

    def step(self, real, gen, render_opts=None):
        """real/gen: EasyDicts of this rank's micro-batch (see loss.accumulate_gradients).  Returns scalar stats."""
        stats = {}
        self.D.requires_grad_(False)
        stats.update(self._phase('Gmain', self.G, self.G_opt, real, gen, 1, render_opts))
        self.G.requires_grad_(False)
        stats.update(self._phase('Dmain', self.D, self.D_opt, real, gen, 1, render_opts))
        if self.D_reg_interval and self.it % self.D_reg_interval == 0:
            stats.update(self._phase('Dreg', self.D, self.D_opt, real, gen, self.D_reg_interval, render_opts))
        # G_ema (training_loop.py:357-366)
        bs = self.batch_size or (len(gen.z) * self.world_size)
        ema_nimg = self.ema_kimg * 1000
        if self.ema_rampup is not None:
            ema_nimg = min(ema_nimg, self.cur_nimg * self.ema_rampup)
        beta = 0.5 ** (bs / max(ema_nimg, 1e-8))
        with torch.no_grad():
            for pe, p in zip(self.G_ema.parameters(), self.G.parameters()):
                pe.copy_(p.lerp(pe, beta))
            for be, b in zip(self.G_ema.buffers(), self.G.buffers()):
                be.copy_(b)
        self.cur_nimg += bs
        self.it += 1
        return stats
