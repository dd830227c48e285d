"""Patch discriminator with hyper-modulation and scalar-encoded patch parameters
(reference src/training/networks_discriminator.py), same classes / arguments / state-dict keys."""
import numpy as np
import torch

from ..torch_utils.ops import conv2d_gradfix, upfirdn2d
from .layers import Conv2dLayer, FullyConnectedLayer, MappingNetwork, ScalarEncoder1d


LOW_PRECISION_TERMS = 16     # ops.tc.operand_formats: 16 = fp16 activations / weights + bf16 gradients; 1 = bf16 everywhere


class DiscriminatorBlock(torch.nn.Module):
    def __init__(self, cfg, in_channels, tmp_channels, out_channels, resolution, img_channels, first_layer_idx, activation='lrelu',
                 resample_filter=[1, 3, 3, 1], conv_clamp=None, use_fp16=False, fp16_channels_last=False, freeze_layers=0, down=2,
                 c_dim=0, hyper_mod=False):
        assert in_channels in [0, tmp_channels]
        super().__init__()
        self.cfg, self.in_channels, self.resolution, self.img_channels = cfg, in_channels, resolution, img_channels
        self.first_layer_idx, self.use_fp16 = first_layer_idx, use_fp16
        self.channels_last = (use_fp16 and fp16_channels_last)
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
        self.num_layers = 0

        def trainable():
            t = (self.first_layer_idx + self.num_layers) >= freeze_layers
            self.num_layers += 1
            return t
        cl = self.channels_last
        self.fromrgb = Conv2dLayer(img_channels, tmp_channels, kernel_size=1, activation=activation, c_dim=c_dim, hyper_mod=False,
                                   trainable=trainable(), conv_clamp=conv_clamp, channels_last=cl)
        self.conv0 = Conv2dLayer(tmp_channels, tmp_channels, kernel_size=3, activation=activation, c_dim=c_dim, hyper_mod=False,
                                 trainable=trainable(), conv_clamp=conv_clamp, channels_last=cl)
        self.conv1 = Conv2dLayer(tmp_channels, out_channels, kernel_size=3, activation=activation, down=down, c_dim=c_dim, hyper_mod=hyper_mod,
                                 trainable=trainable(), resample_filter=resample_filter, conv_clamp=conv_clamp, channels_last=cl)
        self.skip = Conv2dLayer(tmp_channels, out_channels, kernel_size=1, bias=False, down=down, c_dim=c_dim, hyper_mod=False,
                                trainable=trainable(), resample_filter=resample_filter, channels_last=cl)

    def forward(self, x, img, c=None, force_fp32=False):
        """Blocks the reference runs in fp16 (res >= 32, networks_discriminator.py:240) keep float32 STORAGE here and issue their convolutions as
        single-product tensor-core convs with fp16 activation / weight operands (bf16 for gradients) and fp32 accumulation
        (conv2d_gradfix.tc_terms(LOW_PRECISION_TERMS)): the reference's arithmetic class (fp16 operands, fp32 accumulate) without its fp16 storage
        rounding; the fp32 blocks use the error-compensated bf16x3 form.  This removes every fp16<->fp32 cast pass of the reference path."""
        low_precision = self.use_fp16 and not force_fp32
        if x is not None:
            x = x.to(dtype=torch.float32)
        with conv2d_gradfix.tc_terms(LOW_PRECISION_TERMS if low_precision else 3):
            if self.in_channels == 0:
                y = self.fromrgb(img.to(dtype=torch.float32), c=c)
                x = x + y if x is not None else y
            y = self.skip(x, c=c, gain=np.sqrt(0.5))
            x = self.conv0(x, c=c)
            x = self.conv1(x, c=c, gain=np.sqrt(0.5))
        return y.add_(x)


class MinibatchStdLayer(torch.nn.Module):
    def __init__(self, group_size, num_channels=1):
        super().__init__()
        self.group_size, self.num_channels = group_size, num_channels

    def forward(self, x):
        N, C, H, W = x.shape
        G = min(self.group_size, N) if self.group_size is not None else N
        Fc = self.num_channels
        y = x.reshape(G, -1, Fc, C // Fc, H, W)
        y = y - y.mean(dim=0)
        y = (y.square().mean(dim=0) + 1e-8).sqrt()
        y = y.mean(dim=[2, 3, 4]).reshape(-1, Fc, 1, 1).repeat(G, 1, H, W)
        return torch.cat([x, y], dim=1)


class DiscriminatorEpilogue(torch.nn.Module):
    def __init__(self, in_channels, cmap_dim, resolution, img_channels, mbstd_group_size=4, mbstd_num_channels=1, activation='lrelu',
                 conv_clamp=None, feat_predict_dim=0):
        super().__init__()
        self.in_channels, self.cmap_dim, self.resolution, self.img_channels = in_channels, cmap_dim, resolution, img_channels
        self.mbstd = MinibatchStdLayer(group_size=mbstd_group_size, num_channels=mbstd_num_channels) if mbstd_num_channels > 0 else None
        self.conv = Conv2dLayer(in_channels + mbstd_num_channels, in_channels, kernel_size=3, activation=activation, conv_clamp=conv_clamp)
        self.fc = FullyConnectedLayer(in_channels * (resolution ** 2), out_features=in_channels, activation=activation)
        self.out = FullyConnectedLayer(in_channels, out_features=(1 if cmap_dim == 0 else cmap_dim))
        self.feat_out = torch.nn.Sequential(
            FullyConnectedLayer(in_channels * (resolution ** 2), out_features=in_channels, activation=activation),
            FullyConnectedLayer(in_channels, feat_predict_dim)) if feat_predict_dim > 0 else None

    def forward(self, x, cmap, force_fp32=False, predict_feat=False):
        x = x.to(dtype=torch.float32, memory_format=torch.contiguous_format)
        if self.mbstd is not None:
            x = self.mbstd(x)
        x = self.conv(x).flatten(1)
        f = self.feat_out(x) if predict_feat else None
        x = self.out(self.fc(x))
        if self.cmap_dim > 0:
            x = (x * cmap).sum(dim=1, keepdim=True) * (1 / np.sqrt(self.cmap_dim))
        return x, f


class Discriminator(torch.nn.Module):
    def __init__(self, cfg, input_resolution, img_channels, num_fp16_res=4, conv_clamp=256, cmap_dim=None, block_kwargs={},
                 mapping_kwargs={}, epilogue_kwargs={}):
        super().__init__()
        self.cfg = cfg
        assert cfg.num_additional_start_blocks >= 0
        self.img_resolution = input_resolution * (2 ** cfg.num_additional_start_blocks)
        self.img_resolution_log2 = int(np.log2(self.img_resolution))
        self.block_resolutions = [2 ** i for i in range(self.img_resolution_log2, 2, -1)]
        self.img_channels = img_channels
        ch = {res: min(int(cfg.cbase * cfg.fmaps) // res, cfg.cmax) for res in self.block_resolutions + [4]}
        fp16_res = max(2 ** (self.img_resolution_log2 + 1 - num_fp16_res), 8)
        if cmap_dim is None:
            cmap_dim = ch[4]
        self.scalar_enc = ScalarEncoder1d(coord_dim=3, x_multiplier=1000.0, const_emb_dim=256) if cfg.patch.patch_params_cond > 0 else None
        if cfg.c_dim == 0 and self.scalar_enc is None and not cfg.camera_cond:
            cmap_dim = 0
        if cfg.hyper_mod:
            hyper_dim = 512
            self.hyper_mod_mapping = MappingNetwork(z_dim=0, c_dim=self.scalar_enc.get_dim(), camera_cond=False, camera_cond_drop_p=0.0,
                                                    w_dim=hyper_dim, num_ws=None, w_avg_beta=None, **mapping_kwargs)
        else:
            self.hyper_mod_mapping, hyper_dim = None, 0
        common = dict(img_channels=img_channels, conv_clamp=conv_clamp)
        total_c = cfg.c_dim + (0 if self.scalar_enc is None else self.scalar_enc.get_dim())
        idx = 0
        for i, res in enumerate(self.block_resolutions):
            block = DiscriminatorBlock(cfg, ch[res] if res < self.img_resolution else 0, ch[res], ch[res // 2], resolution=res,
                                       first_layer_idx=idx, use_fp16=(res >= fp16_res), down=(1 if i < cfg.num_additional_start_blocks else 2),
                                       c_dim=hyper_dim, hyper_mod=cfg.hyper_mod, **block_kwargs, **common)
            setattr(self, f'b{res}', block)
            idx += block.num_layers
        if cfg.c_dim > 0 or self.scalar_enc is not None:
            self.head_mapping = MappingNetwork(z_dim=0, c_dim=total_c, camera_cond=cfg.camera_cond, camera_cond_drop_p=cfg.camera_cond_drop_p,
                                               w_dim=cmap_dim, num_ws=None, w_avg_beta=None, **mapping_kwargs)
        else:
            self.head_mapping = None
        self.b4 = DiscriminatorEpilogue(ch[4], cmap_dim=cmap_dim, resolution=4, **epilogue_kwargs, **common)

    def forward(self, img, c, patch_params=None, camera_angles=None, update_emas=False, predict_feat=False, **block_kwargs):
        enc = None
        if self.scalar_enc is not None:
            ppc = torch.cat([patch_params['scales'][:, [0]], patch_params['offsets']], dim=1)
            enc = self.scalar_enc(ppc)
            c = torch.cat([c, enc], dim=1)
        hyper_c = self.hyper_mod_mapping(z=None, c=enc) if self.hyper_mod_mapping is not None else None
        x = None
        for res in self.block_resolutions:
            x = getattr(self, f'b{res}')(x, img, c=hyper_c, **block_kwargs)
        cmap = self.head_mapping(z=None, c=c, camera_angles=camera_angles) if self.head_mapping is not None else None
        x, f = self.b4(x, cmap, predict_feat=predict_feat)
        return x.squeeze(1), f
