"""FC / mapping / conv building blocks with the reference's parameterisation and state-dict keys
(reference src/training/layers.py).  Activations and resampling run on lib3dgp_b200's kernels through
torch_utils.ops; dense contractions go through conv2d_gradfix / torch.addmm."""
import contextlib

import numpy as np
import torch
import torch.nn.functional as F

from ..torch_utils.ops import tc, bias_act, conv2d_gradfix, conv2d_resample, modconv, upfirdn2d


def normalize_2nd_moment(x, dim=1, eps=1e-8):
    return x * (x.square().mean(dim=dim, keepdim=True) + eps).rsqrt()


class FullyConnectedLayer(torch.nn.Module):
    """y = act(x @ (W * lr_mul / sqrt(fan_in))^T + b * lr_mul)  (layers.py:22-61)."""

    def __init__(self, in_features, out_features, activation='linear', bias=True, lr_multiplier=1, weight_init=1, bias_init=0):
        super().__init__()
        self.in_features, self.out_features, self.activation = in_features, out_features, activation
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) * (weight_init / lr_multiplier))
        b0 = np.broadcast_to(np.asarray(bias_init, dtype=np.float32), [out_features])
        self.bias = torch.nn.Parameter(torch.from_numpy(b0 / lr_multiplier)) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x):
        w = self.weight.to(x.dtype) * self.weight_gain
        b = self.bias
        if b is not None:
            b = b.to(x.dtype)
            if self.bias_gain != 1:
                b = b * self.bias_gain
        if self.activation == 'linear' and b is not None:
            return torch.addmm(b.unsqueeze(0), x, w.t())
        return bias_act.bias_act(x.matmul(w.t()), b, act=self.activation)

    def extra_repr(self):
        return f'in_features={self.in_features:d}, out_features={self.out_features:d}, activation={self.activation:s}'


class FourierEncoder1d(torch.nn.Module):
    """sin/cos features at log-spaced frequencies (layers.py:304-350)."""

    def __init__(self, coord_dim, max_x_value=100.0, use_cos=True):
        super().__init__()
        self.coord_dim, self.use_cos = coord_dim, use_cos
        n = int(np.ceil(np.log2(max_x_value)))
        coefs = (torch.tensor([2.0]).repeat(n) ** torch.arange(n) / (2 ** n)).float() * np.pi
        self.register_buffer('fourier_coefs', coefs)
        self.fourier_dim = n

    def get_dim(self):
        return self.fourier_dim * (2 if self.use_cos else 1)

    def forward(self, x):
        raw = self.fourier_coefs.view(1, 1, -1) * x.float().unsqueeze(2)
        return torch.cat([raw.sin(), raw.cos()], dim=2) if self.use_cos else raw.sin()


class ScalarEncoder1d(torch.nn.Module):
    """Encodes scalars in [0,1] with Fourier features + a learned table on the rounded value (layers.py:251-299)."""

    def __init__(self, coord_dim, x_multiplier, const_emb_dim, use_raw=False):
        super().__init__()
        self.coord_dim, self.const_emb_dim, self.x_multiplier, self.use_raw = coord_dim, const_emb_dim, x_multiplier, use_raw
        self.const_embed = torch.nn.Embedding(int(np.ceil(x_multiplier)) + 1, const_emb_dim) if (const_emb_dim > 0 and x_multiplier > 0) else None
        self.fourier_encoder = FourierEncoder1d(coord_dim, max_x_value=x_multiplier) if x_multiplier > 0 else None
        self.fourier_dim = self.fourier_encoder.get_dim() if self.fourier_encoder is not None else 0
        self.raw_dim = 1 if use_raw else 0

    def get_dim(self):
        return self.coord_dim * (self.const_emb_dim + self.fourier_dim + self.raw_dim)

    def forward(self, x):
        B = x.shape[0]
        parts = []
        if self.use_raw:
            parts.append(x.unsqueeze(2))
        if self.fourier_encoder is not None or self.const_embed is not None:
            x = x.float() * self.x_multiplier
        if self.fourier_encoder is not None:
            parts.append(self.fourier_encoder(x))
        if self.const_embed is not None:
            parts.append(self.const_embed(x.round().long()))
        return torch.cat(parts, dim=2).view(B, -1)


class MappingNetwork(torch.nn.Module):
    """z, c -> w (layers.py:66-177).  Camera conditioning (camera_cond) is supported as in the reference."""

    def __init__(self, z_dim, c_dim, w_dim, num_ws, num_layers=2, embed_features=None, layer_features=None, activation='lrelu',
                 lr_multiplier=0.01, w_avg_beta=0.998, camera_cond=False, camera_cond_drop_p=0.0, camera_raw_scalars=False,
                 mean_camera_params=None):
        super().__init__()
        if camera_cond:
            self.camera_scalar_enc = (ScalarEncoder1d(coord_dim=2, x_multiplier=0.0, const_emb_dim=0, use_raw=True) if camera_raw_scalars
                                      else ScalarEncoder1d(coord_dim=2, x_multiplier=64.0, const_emb_dim=0))
            c_dim = c_dim + self.camera_scalar_enc.get_dim()
        else:
            self.camera_scalar_enc = None
        self.z_dim, self.c_dim, self.w_dim, self.num_ws, self.num_layers = z_dim, c_dim, w_dim, num_ws, num_layers
        self.w_avg_beta, self.camera_cond_drop_p = w_avg_beta, camera_cond_drop_p
        if self.c_dim > 0:
            embed_features = w_dim if embed_features is None else embed_features
            self.embed = FullyConnectedLayer(self.c_dim, embed_features)
        else:
            embed_features = 0
        layer_features = w_dim if layer_features is None else layer_features
        feats = [z_dim + embed_features] + [layer_features] * (num_layers - 1) + [w_dim]
        for i in range(num_layers):
            setattr(self, f'fc{i}', FullyConnectedLayer(feats[i], feats[i + 1], activation=activation, lr_multiplier=lr_multiplier))
        if num_ws is not None and w_avg_beta is not None:
            self.register_buffer('w_avg', torch.zeros([w_dim]))
        if mean_camera_params is not None:
            self.register_buffer('mean_camera_params', mean_camera_params)
        else:
            self.mean_camera_params = None

    def forward(self, z, c, camera_angles=None, truncation_psi=1, truncation_cutoff=None, update_emas=False):
        if self.camera_scalar_enc is not None:
            if (not self.training) and camera_angles is None:
                camera_angles = self.mean_camera_params[:3].unsqueeze(0).repeat(len(z), 1)
            a = camera_angles[:, [0, 1]]
            a = a.sign() * ((a.abs() % (2.0 * np.pi)) / (2.0 * np.pi))
            emb = F.dropout(self.camera_scalar_enc(a), p=self.camera_cond_drop_p, training=self.training)
            c = torch.cat([torch.zeros(len(emb), 0, device=emb.device) if c is None else c, emb], dim=1)
        x = None
        if self.z_dim > 0:
            x = normalize_2nd_moment(z.to(torch.float32))
        if self.c_dim > 0:
            y = normalize_2nd_moment(self.embed(c.to(torch.float32)))
            x = torch.cat([x, y], dim=1) if x is not None else y
        for i in range(self.num_layers):
            x = getattr(self, f'fc{i}')(x)
        if update_emas and self.w_avg_beta is not None:
            self.w_avg.copy_(x.detach().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        if self.num_ws is not None:
            x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            assert self.w_avg_beta is not None
            if self.num_ws is None or truncation_cutoff is None:
                x = self.w_avg.lerp(x, truncation_psi)
            else:
                x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x


_first_order = True     # Conv2dLayer may use first-order-only fused nodes (ops/modconv.py::_ConvBiasAct, _ChannelScale); see first_order_only()


@contextlib.contextmanager
def first_order_only(flag=True):
    """Scope inside which Conv2dLayer forwards are (flag=True, the default state) / are not (flag=False) allowed to run as first-order-only fused
    nodes.  Code that differentiates the discriminator TWICE (R1, loss.py:316-327; any gradient penalty) wraps its forward in
    `first_order_only(False)`: that forward then stays on the twice-differentiable composition (conv2d_gradfix + bias_act)."""
    global _first_order
    old = _first_order
    _first_order = bool(flag)
    try:
        yield
    finally:
        _first_order = old


class _ChannelScale(torch.autograd.Function):
    """y = x * s[n, c] (hyper-modulation of Conv2dLayer, layers.py:231-232) with a one-pass backward: dx = dy * s and
    g_s[n, c] = sum_hw dy * x come out of ONE kernel (gp3d_modulate_bwd) instead of mul + mul + sum.  First-order only."""

    @staticmethod
    def forward(ctx, x, s):
        ctx.save_for_backward(x, s)
        return x * s.to(x.dtype).unsqueeze(2).unsqueeze(3)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, s = ctx.saved_tensors
        N, C, H, W = x.shape
        if (x.dtype == torch.float32 and dy.dtype == torch.float32 and C % 4 == 0 and x.stride(1) == 1 and x.stride(3) == C and x.stride(2) == W * C
                and x.stride(0) == H * W * C and (x.data_ptr() % 16 == 0)):
            from .. import _lib
            dyn = dy.contiguous(memory_format=torch.channels_last)
            st = s.to(torch.float32).contiguous()
            dx = torch.empty_like(x)                      # preserves the channel-minor strides
            g_s = torch.zeros_like(st)
            with torch.cuda.device(x.device):
                rc = _lib.lib().gp3d_modulate_bwd(dyn.data_ptr(), x.data_ptr(), st.data_ptr(), dx.data_ptr(), g_s.data_ptr(), N, H * W, C, _lib.stream_ptr())
            _lib.check(rc, 'modulate_bwd')
            return dx, g_s.to(s.dtype)
        sb = s.to(x.dtype).unsqueeze(2).unsqueeze(3)
        return dy * sb, (dy.to(torch.float32) * x.to(torch.float32)).sum([2, 3]).to(s.dtype)


class Conv2dLayer(torch.nn.Module):
    """conv (+FIR resampling) + bias + activation, optional hyper-modulation x * (1 + tanh(affine(c)))
    (layers.py:182-246)."""

    def __init__(self, in_channels, out_channels, kernel_size, bias=True, activation='linear', up=1, down=1, resample_filter=[1, 3, 3, 1],
                 conv_clamp=None, channels_last=False, trainable=True, c_dim=0, hyper_mod=False):
        super().__init__()
        self.in_channels, self.out_channels, self.activation, self.up, self.down, self.conv_clamp = in_channels, out_channels, activation, up, down, conv_clamp
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
        self.padding = kernel_size // 2
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))
        self.act_gain = bias_act.activation_funcs[activation].def_gain
        mf = torch.channels_last if channels_last else torch.contiguous_format
        weight = torch.randn([out_channels, in_channels, kernel_size, kernel_size]).to(memory_format=mf)
        b = torch.zeros([out_channels]) if bias else None
        if trainable:
            self.weight = torch.nn.Parameter(weight)
            self.bias = torch.nn.Parameter(b) if b is not None else None
        else:
            self.register_buffer('weight', weight)
            if b is not None:
                self.register_buffer('bias', b)
            else:
                self.bias = None
        self.affine = FullyConnectedLayer(c_dim, in_channels, bias_init=0) if hyper_mod else None
        if hyper_mod:
            assert c_dim > 0

    def forward(self, x, c=None, gain=1):
        k = self.weight.shape[2]
        if (_first_order and isinstance(self.weight, torch.nn.Parameter)
                and modconv.conv_act_eligible(x, self.weight, k, self.up, self.down, self.padding, self.activation, self.in_channels, self.out_channels)):
            # first-order fused node (ops/modconv.py::_ConvBiasAct): hyper-modulation in the operand split, bias / activation / gain / clamp in the
            # conv epilogue; the loss switches this off around the R1 forward, which must stay twice differentiable
            s = (1.0 + self.affine(c).tanh()) if self.affine is not None else None
            alpha = bias_act.activation_funcs[self.activation].def_alpha
            return modconv.conv_bias_act(x, self.weight, self.bias, s, self.weight_gain, self.activation, alpha if alpha is not None else 0.0,
                                         self.act_gain * gain, self.conv_clamp * gain if self.conv_clamp is not None else None,
                                         conv2d_gradfix._terms_for(x.dtype))
        w = self.weight * self.weight_gain
        if self.affine is not None:
            if _first_order and x.is_cuda:
                x = _ChannelScale.apply(x, 1.0 + self.affine(c).tanh())
            else:
                x = (x * (1.0 + self.affine(c).tanh().unsqueeze(2).unsqueeze(3)).to(x.dtype)).to(x.dtype)
        w = w.to(x.dtype)
        if x.is_cuda and isinstance(self.weight, torch.nn.Parameter):
            tc.tag_weight_source(w, self.weight, self.weight_gain)     # lets the tensor-core wrappers reuse this parameter's bf16 operands
        x = conv2d_resample.conv2d_resample(x=x, w=w, f=self.resample_filter, up=self.up, down=self.down,
                                            padding=self.padding, flip_weight=(self.up == 1))
        act_clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        b = self.bias.to(x.dtype) if self.bias is not None else None
        return bias_act.bias_act(x, b, act=self.activation, gain=self.act_gain * gain, clamp=act_clamp)
