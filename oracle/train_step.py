"""TEST / BASELINE INFRASTRUCTURE ONLY -- one EXECUTED optimisation step of the reference algorithm on the host cores.

This is the CPU arm of bench.py (`cpu_baseline`, `--impl reference`): the restated generator / renderer / discriminator of
oracle/restated.py with `DIFFERENTIABLE = True` (every op a torch-CPU op, i.e. exactly the arithmetic the reference's own CPU path
runs: native-PyTorch fallbacks of torch_utils.ops, forced fp32, ATen convolutions), driven through the phases of the reference
training loop:

    Gmain : G forward -> D forward -> softplus(-logits).mean()            -> backward into G      (loss.py:97-113)
    Dmain : G forward (no grad) -> D(fake) ; D(real patch) + KD           -> backward into D      (loss.py:256-316)
    Dreg  : lazy R1 on the real patch, every `d_reg_interval`-th step     -> double backward      (loss.py:316-327)
    Adam (beta1 = 0) on G after Gmain and on D after Dmain / Dreg, G_ema lerp                     (training_loop.py:190-205, 333-366)

Nothing is extrapolated: `CpuTrainer.step()` runs the forward AND backward passes and the optimiser updates it is timed for.
The product package never imports this module.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import restated as R


def extract_patches(x, patch_scales, patch_offsets, resolution):
    """training/training_utils.py:22-53: bilinear patch crop of the real image (grid_sample, align_corners=True)."""
    B = x.shape[0]
    lin = torch.linspace(-1, 1, resolution)
    gx, gy = torch.meshgrid(lin, lin, indexing='xy')
    coords = torch.stack([gx, -gy], dim=-1).unsqueeze(0).repeat(B, 1, 1, 1)       # generate_coords: x right, y up
    coords = (coords + 1.0) * patch_scales.view(B, 1, 1, 2) - 1.0 + patch_offsets.view(B, 1, 1, 2) * 2.0
    coords = torch.stack([coords[..., 0], -coords[..., 1]], dim=-1)
    return F.grid_sample(x, coords, mode='bilinear', align_corners=True)


class CpuTrainer:
    """Holds leaf parameter dicts for G and D (state-dict keyed, as oracle/restated.py consumes them) and two Adam optimisers."""

    def __init__(self, sdG, sdD, Gc, Dc, meta, d_reg_interval=16, r1_gamma=1.0, kd_weight=1.0, seed=0):
        R.DIFFERENTIABLE = True
        R.FAST_FIR = True
        self.Gc, self.Dc, self.meta = Gc, Dc, meta
        buf = ('resample_filter', 'fourier_coefs', 'progress_coef', 'w_avg', 'noise_const')
        self.sdG = {k: (v.clone().requires_grad_(not k.endswith(buf))) for k, v in sdG.items()}
        self.sdD = {k: (v.clone().requires_grad_(not k.endswith(buf))) for k, v in sdD.items()}
        self.pG = [v for v in self.sdG.values() if v.requires_grad]
        self.pD = [v for v in self.sdD.values() if v.requires_grad]
        mb = d_reg_interval / (d_reg_interval + 1) if d_reg_interval else 1.0
        self.optG = torch.optim.Adam(self.pG, lr=0.0025, betas=(0.0, 0.99), eps=1e-8)
        self.optD = torch.optim.Adam(self.pD, lr=0.002 * mb, betas=(0.0 ** mb, 0.99 ** mb), eps=1e-8)
        self.ema = [p.detach().clone() for p in self.pG]
        self.d_reg_interval, self.r1_gamma, self.kd_weight = d_reg_interval, r1_gamma, kd_weight
        self.it = 0
        self.gen = torch.Generator().manual_seed(seed)
        tri = Gc['tri_plane']
        self.block_res = [2 ** i for i in range(2, int(math.log2(tri['res'])) + 1)]
        top = meta['patch_res'] * 2 ** Dc['num_additional_start_blocks']
        self.d_res = [2 ** i for i in range(int(math.log2(top)), 2, -1)]
        self.num_ws = 2 * len(self.block_res)

    # -- pieces -------------------------------------------------------------------------------------------------
    def _run_G(self, z, c, cam, pp):
        B = z.shape[0]
        pr, N = self.meta['patch_res'], self.Gc['num_ray_steps']
        ws = R.mapping_network(self.sdG, 'mapping.', z, c, self.num_ws)
        noises = [torch.randn(B, 1, r, r, generator=self.gen) for r in self.block_res for _ in range(1 if r == 4 else 2)]
        u1 = torch.rand(B, pr * pr, N, generator=self.gen); u2 = torch.rand(B, pr * pr, N, generator=self.gen)
        heads = torch.randint(0, self.Gc['depth_adaptor']['num_hid_layers'] + 1, (B,), generator=self.gen)
        out = R.generator_synthesis(self.sdG, self.Gc, ws, cam['angles'], cam['fov'], cam['radius'], cam['look_at'], pr, pp[0], pp[1], u1, u2,
                                    noise_mode='random', noises=noises, fused_modconv=False, depth_head_idx=heads)
        return out['img']

    def _run_D(self, img, c, pp, predict_feat=False):
        return R.discriminator(self.sdD, img, c, pp[0], pp[1], self.d_res, self.Dc['num_additional_start_blocks'], predict_feat=predict_feat)

    def _patch_params(self, B):
        s = torch.full((B, 2), float(self.meta['patch_res']) / self.meta['img_resolution'])
        o = torch.rand(1, 2, generator=self.gen).repeat(B, 1) * (1.0 - s)
        return s, o

    def _set_grad(self, params, flag):
        for p in params:
            p.requires_grad_(flag)

    # -- one iteration ------------------------------------------------------------------------------------------
    def step(self, real_img, real_depth, c, embs, z, cam):
        """real_img [B,3,H,W], real_depth [B,1,H,W], c one-hot, embs [B,E], z latents, cam dict(angles, fov, radius, look_at)."""
        B = z.shape[0]
        stats = {}
        # Gmain
        self._set_grad(self.pD, False); self._set_grad(self.pG, True)
        self.optG.zero_grad(set_to_none=True)
        pp = self._patch_params(B)
        logits, _ = self._run_D(self._run_G(z, c, cam, pp), c, pp)
        lossG = F.softplus(-logits).mean()
        lossG.backward()
        for p in self.pG:
            if p.grad is not None:
                torch.nan_to_num(p.grad, nan=0, posinf=1e5, neginf=-1e5, out=p.grad)
        self.optG.step()
        stats['Loss/G/loss'] = float(lossG.detach())
        # Dmain
        self._set_grad(self.pG, False); self._set_grad(self.pD, True)
        self.optD.zero_grad(set_to_none=True)
        with torch.no_grad():
            pp = self._patch_params(B)
            fake = self._run_G(z, c, cam, pp)
        logits_f, _ = self._run_D(fake, c, pp)
        F.softplus(logits_f).mean().backward()
        pp = self._patch_params(B)
        real = extract_patches(torch.cat([real_img, real_depth], dim=1), pp[0], pp[1], self.meta['patch_res'])
        logits_r, feats = self._run_D(real, c, pp, predict_feat=self.kd_weight > 0)
        loss = F.softplus(-logits_r)
        if self.kd_weight > 0:
            loss = loss + (feats - embs).norm(dim=1) * self.kd_weight
        loss.mean().backward()
        self._adam_D()
        stats['Loss/scores/fake'] = float(logits_f.detach().mean()); stats['Loss/scores/real'] = float(logits_r.detach().mean())
        # Dreg (lazy R1)
        if self.d_reg_interval and self.it % self.d_reg_interval == 0:
            self.optD.zero_grad(set_to_none=True)
            pp = self._patch_params(B)
            real = extract_patches(torch.cat([real_img, real_depth], dim=1), pp[0], pp[1], self.meta['patch_res']).detach().requires_grad_(True)
            logits_r, _ = self._run_D(real, c, pp)
            r1 = torch.autograd.grad([logits_r.sum()], [real], create_graph=True)[0]
            pen = r1.square().sum([1, 2, 3])
            (pen * (self.r1_gamma / 2)).mean().mul(self.d_reg_interval).backward()
            self._adam_D()
            stats['Loss/D/r1_penalty'] = float(pen.detach().mean())
        with torch.no_grad():
            for pe, p in zip(self.ema, self.pG):
                pe.copy_(p.detach().lerp(pe, 0.5 ** (B / 10000.0)))
        self.it += 1
        return stats

    def _adam_D(self):
        for p in self.pD:
            if p.grad is not None:
                torch.nan_to_num(p.grad, nan=0, posinf=1e5, neginf=-1e5, out=p.grad)
        self.optD.step()


def random_state_dicts(Gc, Dc, meta, seed=0):
    """Random-init weights with the reference constructors' distributions (randn weights, zero biases, affine bias 1, noise strength 0.1
    so that the noise path is exercised -- SURVEY.md 8d), keyed like the reference state dicts (oracle/shapes.py)."""
    from . import shapes
    gs, _ = shapes.generator_shapes(Gc)
    ds, _ = shapes.discriminator_shapes(Dc, meta['patch_res'], 4, meta['embedding_dim'])
    filt = torch.from_numpy(R.setup_filter([1, 3, 3, 1]))

    def fill(table, s):
        g = torch.Generator().manual_seed(s)
        sd = {}
        for k, shp in table.items():
            if k.endswith('resample_filter'):
                sd[k] = filt.clone()
            elif k.endswith('fourier_coefs'):
                sd[k] = (2.0 ** torch.arange(shp[0]).float() / (2 ** shp[0])) * np.pi
            elif k.endswith('.bias') or k.endswith('progress_coef') or k.endswith('w_avg'):
                sd[k] = torch.zeros(shp) + (1.0 if (k.endswith('affine.bias') and 'synthesis' in k) else 0.0)
            elif k.endswith('noise_strength'):
                sd[k] = torch.tensor(0.1)
            elif k.endswith('near_plane_offset_raw'):
                sd[k] = torch.tensor([-3.0])
            else:
                sd[k] = torch.randn(shp, generator=g)
        return sd
    return fill(gs, seed), fill(ds, seed + 1)
