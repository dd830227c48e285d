"""StyleGAN2 synthesis layers used by the tri-plane decoder (reference src/training/networks_stylegan2.py:31-276),
same classes / constructor arguments / state-dict keys.  Elementwise work runs on lib3dgp_b200 kernels (bias_act,
upfirdn2d); the dense contraction goes through conv2d_resample -> conv2d_gradfix."""
import numpy as np
import torch

from .. import _lib
from ..torch_utils.ops import bias_act, conv2d_resample, fma, modconv, upfirdn2d

fused_layer_enabled = True   # route eligible training-path layers through the fused modulated-conv node (ops/modconv.py)
from .layers import Conv2dLayer, FullyConnectedLayer


def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True):
    """Weight (de)modulation, networks_stylegan2.py:31-88.  fused_modconv=False (training): activations are scaled before and
    after a shared-weight conv; True (inference): per-sample weights through a grouped conv."""
    B = x.shape[0]
    oc, ic, kh, kw = weight.shape
    if x.dtype == torch.float16 and demodulate:     # pre-normalise against fp16 overflow (:51-53)
        weight = weight * (1 / np.sqrt(ic * kh * kw) / weight.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
        styles = styles / styles.norm(float('inf'), dim=1, keepdim=True)
    w = dcoefs = None
    if fused_modconv:
        w = weight.unsqueeze(0) * styles.reshape(B, 1, -1, 1, 1)
    if demodulate and not fused_modconv:
        # sum_{i,k} (w[o,i,k] s[n,i])^2 as a [B,Cin] x [Cin,Cout] product of squares: same value, no [B,Cout,Cin,k,k] temporary
        dcoefs = (styles.square() @ weight.square().sum(dim=[2, 3]).t() + 1e-8).rsqrt()
    elif demodulate:
        dcoefs = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
    if demodulate and fused_modconv:
        w = w * dcoefs.reshape(B, -1, 1, 1, 1)
    if not fused_modconv:
        x = x * styles.to(x.dtype).reshape(B, -1, 1, 1)
        x = conv2d_resample.conv2d_resample(x=x, w=weight.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding, flip_weight=flip_weight)
        if demodulate and noise is not None:
            x = fma.fma(x, dcoefs.to(x.dtype).reshape(B, -1, 1, 1), noise.to(x.dtype))
        elif demodulate:
            x = x * dcoefs.to(x.dtype).reshape(B, -1, 1, 1)
        elif noise is not None:
            x = x.add_(noise.to(x.dtype))
        return x
    x = x.reshape(1, -1, *x.shape[2:])
    w = w.reshape(-1, ic, kh, kw)
    x = conv2d_resample.conv2d_resample(x=x, w=w.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding, groups=B, flip_weight=flip_weight)
    x = x.reshape(B, -1, *x.shape[2:])
    if noise is not None:
        x = x.add_(noise)
    return x


class SynthesisLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, use_noise=True, activation='lrelu',
                 resample_filter=[1, 3, 3, 1], conv_clamp=None, channels_last=False):
        super().__init__()
        self.in_channels, self.out_channels, self.w_dim, self.resolution = in_channels, out_channels, w_dim, resolution
        self.up, self.use_noise, self.activation, self.conv_clamp = up, use_noise, activation, conv_clamp
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
        self.padding = kernel_size // 2
        self.act_gain = bias_act.activation_funcs[activation].def_gain
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        mf = torch.channels_last if channels_last else torch.contiguous_format
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]).to(memory_format=mf))
        if use_noise:
            self.register_buffer('noise_const', torch.randn([resolution, resolution]))
            self.noise_strength = torch.nn.Parameter(torch.zeros([]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))

    def forward(self, x, w, noise_mode='random', fused_modconv=True, gain=1, noise_in=None):
        """noise_in: optional [B,1,r,r] standard-normal image replacing the torch.randn draw of :134 (parity runs)."""
        assert noise_mode in ['random', 'const', 'none']
        styles = self.affine(w)
        noise = None
        if self.use_noise and noise_mode == 'random':
            if noise_in is None:
                noise_in = torch.randn([x.shape[0], 1, x.shape[2] * self.up, x.shape[3] * self.up], device=x.device)
            noise = noise_in * self.noise_strength
        if self.use_noise and noise_mode == 'const':
            noise = self.noise_const * self.noise_strength
        if (fused_layer_enabled and self.activation == 'lrelu' and modconv.eligible(x, self.weight, self.up, self.conv_clamp)):
            # fp32 CUDA path (training AND inference: scaling activations (:67-76) and scaling weights (:78-88) are the same map, Appendix A): x*styles -> conv(+FIR) -> *dcoefs + noise + bias -> lrelu*gain as one autograd node.
            # dcoefs = rsqrt(sum_{i,k} (w[o,i,k] s[n,i])^2 + 1e-8) (:62) evaluated as a [B,Cin] x [Cin,Cout] product of squares.
            dcoefs = (styles.square() @ self.weight.square().sum(dim=[2, 3]).t() + 1e-8).rsqrt()
            nimg = None
            if self.use_noise and noise_mode == 'random':
                nimg = noise_in
            elif self.use_noise and noise_mode == 'const':
                nimg = self.noise_const
            return modconv.modconv_layer(x, self.weight, styles, dcoefs=dcoefs, noise=nimg, noise_strength=self.noise_strength if nimg is not None else None,
                                         bias=self.bias, up=self.up, fir=self.resample_filter, act='lrelu', alpha=0.2, gain=self.act_gain * gain)
        x = modulated_conv2d(x=x, weight=self.weight, styles=styles, noise=noise, up=self.up, padding=self.padding,
                             resample_filter=self.resample_filter, flip_weight=(self.up == 1), fused_modconv=fused_modconv)
        act_clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        return bias_act.bias_act(x, self.bias.to(x.dtype), act=self.activation, gain=self.act_gain * gain, clamp=act_clamp)


class ToRGBLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1, conv_clamp=None, channels_last=False):
        super().__init__()
        self.in_channels, self.out_channels, self.w_dim, self.conv_clamp = in_channels, out_channels, w_dim, conv_clamp
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        mf = torch.channels_last if channels_last else torch.contiguous_format
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]).to(memory_format=mf))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))

    def forward(self, x, w, fused_modconv=True):
        styles = self.affine(w) * self.weight_gain
        if fused_layer_enabled and modconv.eligible(x, self.weight, 1, self.conv_clamp):
            return modconv.modconv_layer(x, self.weight, styles, bias=self.bias, up=1, act='linear', gain=1.0)
        x = modulated_conv2d(x=x, weight=self.weight, styles=styles, demodulate=False, fused_modconv=fused_modconv)
        return bias_act.bias_act(x, self.bias.to(x.dtype), clamp=self.conv_clamp)


class SynthesisBlock(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture='skip', resample_filter=[1, 3, 3, 1],
                 conv_clamp=256, use_fp16=False, fp16_channels_last=False, fused_modconv_default=True, **layer_kwargs):
        assert architecture in ['orig', 'skip', 'resnet']
        super().__init__()
        self.in_channels, self.w_dim, self.resolution, self.img_channels = in_channels, w_dim, resolution, img_channels
        self.is_last, self.architecture, self.use_fp16 = is_last, architecture, use_fp16
        self.channels_last = (use_fp16 and fp16_channels_last)
        self.fused_modconv_default = fused_modconv_default
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
        self.num_conv = 0
        self.num_torgb = 0
        if in_channels == 0:
            self.const = torch.nn.Parameter(torch.randn([out_channels, resolution, resolution]))
        if in_channels != 0:
            self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim=w_dim, resolution=resolution, up=2, resample_filter=resample_filter,
                                        conv_clamp=conv_clamp, channels_last=self.channels_last, **layer_kwargs)
            self.num_conv += 1
        self.conv1 = SynthesisLayer(out_channels, out_channels, w_dim=w_dim, resolution=resolution, conv_clamp=conv_clamp,
                                    channels_last=self.channels_last, **layer_kwargs)
        self.num_conv += 1
        if is_last or architecture == 'skip':
            self.torgb = ToRGBLayer(out_channels, img_channels, w_dim=w_dim, conv_clamp=conv_clamp, channels_last=self.channels_last)
            self.num_torgb += 1
        if in_channels != 0 and architecture == 'resnet':
            self.skip = Conv2dLayer(in_channels, out_channels, kernel_size=1, bias=False, up=2, resample_filter=resample_filter, channels_last=self.channels_last)

    def forward(self, x, img, ws, force_fp32=False, fused_modconv=None, update_emas=False, layer_noises=None, **layer_kwargs):
        w_iter = iter(ws.unbind(dim=1))
        _lib.require_cuda(ws, 'ws')          # no CPU path
        dtype = torch.float16 if self.use_fp16 and not force_fp32 else torch.float32
        mf = torch.channels_last if self.channels_last and not force_fp32 else torch.contiguous_format
        if fused_modconv is None:
            fused_modconv = self.fused_modconv_default
        if fused_modconv == 'inference_only':
            fused_modconv = (not self.training)
        nz = iter(layer_noises) if layer_noises is not None else None
        nxt = (lambda: next(nz)) if nz is not None else (lambda: None)
        if self.in_channels == 0:
            x = self.const.to(dtype=dtype, memory_format=mf).unsqueeze(0).repeat([ws.shape[0], 1, 1, 1])
        else:
            # keep whatever layout the previous block produced (channel-minor on the tensor-core path): a forced
            # memory_format=contiguous here costs one full-tensor copy per block and another one back inside the next conv
            x = x.to(dtype=dtype, memory_format=torch.channels_last) if self.channels_last and not force_fp32 else x.to(dtype=dtype)
        if self.in_channels == 0:
            x = self.conv1(x, next(w_iter), fused_modconv=fused_modconv, noise_in=nxt(), **layer_kwargs)
        elif self.architecture == 'resnet':
            y = self.skip(x, gain=np.sqrt(0.5))
            x = self.conv0(x, next(w_iter), fused_modconv=fused_modconv, noise_in=nxt(), **layer_kwargs)
            x = self.conv1(x, next(w_iter), fused_modconv=fused_modconv, gain=np.sqrt(0.5), noise_in=nxt(), **layer_kwargs)
            x = y.add_(x)
        else:
            x = self.conv0(x, next(w_iter), fused_modconv=fused_modconv, noise_in=nxt(), **layer_kwargs)
            x = self.conv1(x, next(w_iter), fused_modconv=fused_modconv, noise_in=nxt(), **layer_kwargs)
        if img is not None:
            img = upfirdn2d.upsample2d(img, self.resample_filter)
        if self.is_last or self.architecture == 'skip':
            y = self.torgb(x, next(w_iter), fused_modconv=fused_modconv).to(dtype=torch.float32)
            if img is None:
                # keep the skip image channel-minor from its first appearance: every later upsample2d (C-minor kernel) and the
                # final tri-plane tensor then stay in the layout the fused ray-march gathers from, with no extra copy
                img = y.contiguous(memory_format=torch.channels_last) if (y.shape[1] % 4 == 0 and y.shape[1] > 4) else y
            else:
                img = img.add_(y)
        return x, img


class SynthesisNetwork(torch.nn.Module):
    """The 2-D StyleGAN2 image synthesis network (`model=stylegan2`, reference networks_stylegan2.py:281-340): the same SynthesisBlocks from 4^2 up to the image
    resolution, no volume rendering.  Kept so that the file can replace the reference's for every model it serves; the tri-plane decoder of the 3-D models is
    `networks_epigraf.SynthesisBlocksSequence` over the same blocks."""

    def __init__(self, cfg, img_resolution, img_channels, num_fp16_res=4, **block_kwargs):
        assert img_resolution >= 4 and img_resolution & (img_resolution - 1) == 0
        super().__init__()
        self.cfg, self.img_resolution, self.img_channels, self.num_fp16_res = cfg, img_resolution, img_channels, num_fp16_res
        self.img_resolution_log2 = int(np.log2(img_resolution))
        self.block_resolutions = [2 ** i for i in range(2, self.img_resolution_log2 + 1)]
        width = lambda res: min(int(cfg.cbase * cfg.fmaps) // res, cfg.cmax)
        first_fp16 = max(2 ** (self.img_resolution_log2 + 1 - num_fp16_res), 8)
        self.num_ws = 0
        for res in self.block_resolutions:
            last = (res == img_resolution)
            block = SynthesisBlock(width(res // 2) if res > 4 else 0, width(res), w_dim=cfg.w_dim, resolution=res, img_channels=img_channels, is_last=last,
                                   use_fp16=(res >= first_fp16), architecture=cfg.get('architecture', 'skip'), **block_kwargs)
            self.num_ws += block.num_conv + (block.num_torgb if last else 0)
            setattr(self, f'b{res}', block)

    def forward(self, ws, camera_params=None, patch_params=None, render_opts={}, layer_noises=None, **block_kwargs):
        """camera_params / render_opts are accepted for call compatibility with the 3-D generators (the loss passes them to either);
        layer_noises: optional list of [B,1,r,r] noise images, one per noisy layer in execution order (parity runs)."""
        assert not render_opts.get('concat_depth', False), 'a 2-D generator has no depth to concatenate'
        assert tuple(ws.shape[1:]) == (self.num_ws, self.cfg.w_dim), tuple(ws.shape)
        ws = ws.to(torch.float32)
        x = img = None
        first = 0
        for res in self.block_resolutions:
            block = getattr(self, f'b{res}')
            given = None if layer_noises is None else layer_noises[first:first + block.num_conv]
            x, img = block(x, img, ws.narrow(1, first, block.num_conv + block.num_torgb), layer_noises=given, **block_kwargs)
            first += block.num_conv
        if self.training and patch_params is not None:
            from .training_utils import extract_patches
            img = extract_patches(img, patch_params, resolution=self.cfg.patch.resolution)
        if render_opts.get('return_depth', False):
            from ..dnnlib import TensorGroup
            return TensorGroup(img=img, depth=torch.zeros_like(img))
        return img


class Generator(torch.nn.Module):
    """2-D StyleGAN2 generator (networks_stylegan2.py:346-375): mapping network + SynthesisNetwork."""

    def __init__(self, cfg, img_resolution, img_channels, mapping_kwargs={}, **synthesis_kwargs):
        super().__init__()
        from .layers import MappingNetwork
        self.cfg, self.z_dim, self.c_dim, self.w_dim = cfg, cfg.z_dim, cfg.c_dim, cfg.w_dim
        self.img_resolution, self.img_channels = img_resolution, img_channels
        self.synthesis = SynthesisNetwork(cfg=cfg, img_resolution=img_resolution, img_channels=img_channels, **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=self.z_dim, c_dim=self.c_dim, w_dim=self.w_dim, num_ws=self.num_ws, num_layers=cfg.map_depth, **mapping_kwargs)
        self.params_to_freeze = None

    def forward(self, z, c, camera_angles_cond=None, truncation_psi=1, truncation_cutoff=None, update_emas=False, **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.synthesis(ws, update_emas=update_emas, **synthesis_kwargs)

    def progressive_update(self, cur_kimg):
        pass

