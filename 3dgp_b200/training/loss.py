"""Non-saturating GAN loss with R1, patch-wise training and discriminator knowledge distillation
(reference src/training/loss.py:33-339), `pl_weight=0` (3dgp.yaml / base.yaml:77).
Phases: Gmain, Dmain, Dreg (lazy R1) exactly as training_loop.py:321-331 drives them.
`training.learn_camera_dist=true`: the camera adaptor maps the prior camera before rendering (:78-79) and Gmain adds its regularisers
(:142-232): Lipschitz (off in 3dgp.yaml), earth-mover distance between prior and posterior per camera coordinate, and the pull of the mean
origin angles to the prior mean.  The reference calls POT (`ot.dist` + `ot.emd2`, environment.yml:38, unpinned version) for the EMD; POT is not
installed here, so the term is restated: for two equally weighted 1-D samples of the same size and a convex cost the optimal plan matches
order statistics, hence emd2 == mean((sort(a) - sort(b))^2), differentiated with the plan held fixed exactly as `ot.emd2` does (`emd2_1d`).
POT itself is absent, so the term is pinned against an independent EXACT solver of the same transport problem instead (uniform weights and equal
sample counts make it an assignment problem: scipy's `linear_sum_assignment` on the squared-distance matrix `ot.dist` builds; value and gradient,
tests/test_cpu_oracle.py::test_emd_restatement_equals_the_exact_assignment_solution)."""
import numpy as np
import torch

from . import layers

from ..dnnlib import EasyDict
from ..torch_utils.ops import conv2d_gradfix, upfirdn2d
from .rendering_utils import get_mean_angles_values
from .training_utils import extract_patches, linear_schedule, sample_patch_params, sample_random_c


def maybe_blur(img, blur_sigma):
    """Gaussian blur with exp2 taps over +-3 sigma (loss.py:331-337)."""
    blur_size = np.floor(blur_sigma * 3)
    if blur_size > 0:
        f = torch.arange(-blur_size, blur_size + 1, device=img.device).div(blur_sigma).square().neg().exp2()
        img = upfirdn2d.filter2d(img, f / f.sum())
    return img


def emd2_1d(a, b):
    """Per-column squared-Euclidean earth-mover distance between two equally weighted samples a, b [n, k] -> [1, k]: what
    `ot.emd2(1/n, 1/n, ot.dist(a[:, [i]], b[:, [i]]))` returns for every column i (loss.py:195-197).  On the line the monotone (sorted) matching is
    optimal for a convex cost; autograd through the gather keeps the plan fixed, like POT's gradient of emd2 with respect to the cost matrix."""
    return (a.sort(dim=0).values - b.sort(dim=0).values).square().mean(dim=0, keepdim=True)


class StyleGAN2Loss:
    def __init__(self, cfg, device, G, D, augment_pipe=None, r1_gamma=10, style_mixing_prob=0, pl_batch_shrink=2, pl_decay=0.01):
        """Same constructor keywords as the reference class (loss.py:34), so `training_loop.py:186` can build either one.  `augment_pipe` (ADA, augment.py) is
        applied to the discriminator's input as in loss.py:98-99; style mixing and path-length regularisation are off in configs/model/{3dgp,epigraf}.yaml
        (`style_mixing_prob` is only set for model=stylegan2, train.py:200; `pl_weight: 0`) and are not built."""
        if style_mixing_prob > 0:
            raise NotImplementedError('style mixing (model=stylegan2 only, train.py:200) is not on the 3dgp path')
        if cfg.model.loss_kwargs.get('pl_weight', 0) > 0:
            raise NotImplementedError('path-length regularisation (pl_weight > 0) is not on the 3dgp path')
        self.cfg, self.device, self.G, self.D, self.r1_gamma, self.augment_pipe = cfg, device, G, D, r1_gamma, augment_pipe
        lk = cfg.model.loss_kwargs
        self.blur_init_sigma = lk.get('blur_init_sigma', 0)
        self.blur_fade_kimg = lk.get('blur_fade_kimg', 0)
        self.patch_cfg = EasyDict.init_recursively(cfg.training.patch)
        self.progressive_update(0)

    def progressive_update(self, cur_kimg):
        pc = self.patch_cfg
        if pc.enabled:
            if pc.distribution == 'beta':
                pc.beta = linear_schedule(cur_kimg, pc.beta_val_start, pc.beta_val_end, pc.anneal_kimg)
                pc.min_scale = pc.min_scale_trg
            else:
                pc.min_scale = linear_schedule(cur_kimg, pc.max_scale, pc.min_scale_trg, pc.anneal_kimg)
        kd = self.cfg.model.loss_kwargs.kd.discr
        self.D_kd_weight = linear_schedule(cur_kimg, kd.weight, 0.0, period=kd.anneal_kimg, start_step=0)
        self.learn_camera = bool(self.cfg.training.get('learn_camera_dist', False))
        self.emd_multiplier = linear_schedule(cur_kimg, 0.0, 1.0, period=self.cfg.model.generator.camera_adaptor.emd.anneal_kimg, start_step=0) if self.learn_camera else 0.0

    def run_G(self, z, c, camera_params, update_emas=False, patch_params=None, render_opts=None):
        ws = self.G.mapping(z=z, c=c, update_emas=update_emas)
        if patch_params is None:
            patch_params = sample_patch_params(len(z), self.patch_cfg, device=z.device) if self.patch_cfg.enabled else {}
        kw = dict(patch_params=patch_params) if self.patch_cfg.enabled else {}
        if self.learn_camera:
            camera_params = self.G.synthesis.camera_adaptor(camera_params, z, c)                      # loss.py:78-79
        ro = dict(concat_depth=self.cfg.training.use_depth, return_depth=True)
        ro.update(render_opts or {})
        out = self.G.synthesis(ws, camera_params, update_emas=update_emas, render_opts=ro, **kw)
        out.ws = ws
        out.camera_params = camera_params
        return out, patch_params

    def camera_regularisers(self, stats):
        """Gmain's extra terms for the learned camera distribution (loss.py:142-232); returns their sum."""
        ca, ccfg = self.G.synthesis.camera_adaptor, self.cfg.model.generator.camera_adaptor
        total = 0.0

        def prior_posterior(n):
            z = torch.randn(n, self.G.z_dim, device=self.device)
            c = sample_random_c(n, self.G.c_dim, self.device)
            prior_raw = ca.unroll_camera_params(ca.sample_from_prior(n, device=self.device)).requires_grad_(True)      # [n, 8]
            post_raw = ca.unroll_camera_params(ca(ca.roll_camera_params(prior_raw), z, c))
            return prior_raw, post_raw

        def weighted(t, w):        # t [1, 8] in unrolled order; only yaw, pitch, radius, fov and the look-at triple count (:175, :214)
            g = ca.roll_camera_params(t)
            return (g.angles * w[0])[:, :2].sum() + (g.radius * w[1]).sum() + (g.fov * w[2]).sum() + (g.look_at * w[3]).sum()

        if ccfg.lipschitz_weights.enabled:                                                                             # :143-177
            prior_raw, post_raw = prior_posterior(256)
            grads = torch.stack([torch.autograd.grad(post_raw[:, i].sum(), prior_raw, create_graph=True)[0][:, i] for i in range(8)], dim=1).abs()
            lw = ccfg.lipschitz_weights
            lip = weighted((grads + 1.0 / (grads + 1e-4)).mean(dim=0, keepdim=True), (lw.angles, lw.radius, lw.fov, lw.look_at))
            stats['Loss/camera_dist/lipschitz_loss'] = lip.detach()
            total = total + lip
        if ccfg.emd.enabled and self.emd_multiplier > 0.0:                                                             # :182-216
            prior_raw, post_raw = prior_posterior(ccfg.emd.num_samples)
            emd = emd2_1d(post_raw, prior_raw.detach())   # [1, 8]; the prior sample is a leaf without parameters behind it
            emd = self.emd_multiplier * weighted(emd, (ccfg.emd.origin, ccfg.emd.radius, ccfg.emd.fov, ccfg.emd.look_at))
            stats['Loss/camera_dist/emd_loss'] = emd.detach()
            total = total + emd
        if ccfg.adjust.angles and ccfg.force_mean_weight > 0:                                                          # :221-230
            mean_angles = torch.tensor(get_mean_angles_values(self.cfg.camera.origin.angles), device=self.device)
            _, post_raw = prior_posterior(256)
            fm = ccfg.force_mean_weight * (post_raw[:, :3].mean(dim=0) - mean_angles + 1e-8).square().sum().sqrt()
            stats['Loss/camera_dist/force_mean'] = fm.detach()
            total = total + fm
        return total

    def run_D(self, img, c, blur_sigma=0, update_emas=False, **kwargs):
        img = maybe_blur(img, blur_sigma)
        if self.cfg.training.use_depth:
            bs = np.floor(blur_sigma * 3)
            f = torch.arange(-bs, bs + 1, device=img.device).div(30.0).square().neg().exp2()   # loss.py:93-94
            img = torch.cat([img[:, :3], upfirdn2d.filter2d(img[:, [3]], f / f.sum()), img[:, 4:]], dim=1)
        if self.augment_pipe is not None:
            img = self.augment_pipe(img, num_color_channels=self.G.img_channels)
        return self.D(img, c, update_emas=update_emas, **kwargs)

    def compute_sample_weights(self, patch_params, scale_pow=1):
        s = patch_params['scales'].mean(dim=1) ** scale_pow
        return s / (s.mean(dim=0) + 1e-8)

    def accumulate_gradients(self, phase, real_data, gen_data, gain, cur_nimg, render_opts=None, final_backward=None):
        """real_data: {img [B,3,H,W], depth [B,1,H,W], c, embs, camera_angles}; gen_data: {z, c, camera_params}.
        final_backward: optional callable invoked right before the LAST backward() of the phase (training/step.py arms the bucketed all-reduce there)."""
        arm = final_backward if final_backward is not None else (lambda: None)
        phase = {'Gall': 'Gmain'}.get(phase, phase)      # without lazy G regularisation the reference loop names G's only phase 'Gall' (training_loop.py:195); pl_weight is 0 here
        assert phase in ['Gmain', 'Dmain', 'Dreg', 'Dall']
        if self.r1_gamma == 0:
            phase = {'Dreg': 'none', 'Dall': 'Dmain'}.get(phase, phase)
        blur_sigma = max(1 - cur_nimg / (self.blur_fade_kimg * 1e3), 0) * self.blur_init_sigma if self.blur_fade_kimg > 0 else 0
        lk = self.cfg.model
        stats = {}
        real_img = torch.cat([real_data.img, real_data.depth], dim=1) if self.cfg.training.use_depth else real_data.img

        if phase == 'Gmain':
            gen_out, pp = self.run_G(gen_data.z, gen_data.c, gen_data.camera_params, render_opts=render_opts)
            logits, _ = self.run_D(gen_out.img, gen_data.c, blur_sigma=blur_sigma, patch_params=pp, camera_angles=gen_out.camera_params.angles)
            loss = torch.nn.functional.softplus(-logits)
            reg = self.camera_regularisers(stats) if self.learn_camera else 0.0
            arm()
            (loss.mean() + reg).mul(gain).backward()
            stats['Loss/G/loss'] = loss.detach().mean()

        loss_Dgen = 0
        if phase in ['Dmain', 'Dall']:
            with torch.no_grad():
                gen_out, pp = self.run_G(gen_data.z, gen_data.c, gen_data.camera_params, update_emas=True, render_opts=render_opts)
            logits, _ = self.run_D(gen_out.img, gen_data.c, blur_sigma=blur_sigma, update_emas=True, patch_params=pp, camera_angles=gen_out.camera_params.angles)
            loss_Dgen = torch.nn.functional.softplus(logits.clamp(min=-lk.discriminator.logits_clamp_val))
            loss_Dgen = loss_Dgen + 0.0 * logits.max()
            loss_Dgen.mean().mul(gain).backward()
            stats['Loss/scores/fake'] = logits.detach().mean()

        if phase in ['Dmain', 'Dreg', 'Dall']:
            do_kd = self.D_kd_weight > 0 and phase in ['Dmain', 'Dall'] and self.D.b4.feat_out is not None
            if self.patch_cfg.enabled:
                pp = sample_patch_params(len(real_img), self.patch_cfg, device=real_img.device)
                real_patch = extract_patches(real_img, pp, resolution=self.patch_cfg.resolution)
            else:
                pp, real_patch = None, real_img
            tmp = real_patch.detach().requires_grad_(phase in ['Dreg', 'Dall'])
            with layers.first_order_only(phase not in ['Dreg', 'Dall']):      # R1 differentiates D twice: keep that forward on twice-differentiable ops
                logits, feats = self.run_D(tmp, real_data.c, blur_sigma=blur_sigma, patch_params=pp, camera_angles=real_data.get('camera_angles'), predict_feat=do_kd)
            loss_Dreal = loss_Dkd = loss_Dr1 = 0
            if phase in ['Dmain', 'Dall']:
                loss_Dreal = torch.nn.functional.softplus(-logits.clamp(max=lk.discriminator.logits_clamp_val)) + 0.0 * logits.max()
                stats['Loss/scores/real'] = logits.detach().mean()
            if do_kd:
                dist = (feats - real_data.embs).norm(dim=1) * self.compute_sample_weights(pp)
                loss_Dkd = dist * self.D_kd_weight
            if phase in ['Dreg', 'Dall']:
                with conv2d_gradfix.no_weight_gradients():
                    r1_grads = torch.autograd.grad(outputs=[logits.sum()], inputs=[tmp], create_graph=True, only_inputs=True)[0]
                r1_penalty = r1_grads.square().sum([1, 2, 3])
                loss_Dr1 = r1_penalty * (self.r1_gamma / 2)
                stats['Loss/D/r1_penalty'] = r1_penalty.detach().mean()
            arm()
            (loss_Dreal + loss_Dr1 + loss_Dkd).mean().mul(gain).backward()
        return stats
