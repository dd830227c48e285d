"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy / torch-CPU fp32) of the reference algorithm for 3DGP's
per-image hot path.  It is the parity ORACLE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it.  The product package (3dgp_b200/) never does.

Pinning: the reference (snap-research/3dgp @ /root/reference) ships no tests, golden vectors or fixtures
(SURVEY.md 4, 8c).  This restatement is therefore pinned against the reference ITSELF, imported and executed in the
build container by oracle/make_golden.py (functions `check_*`), and the resulting input/output vectors are committed
under tests/golden/ so that the pin travels to the GPU box where /root/reference does not exist.
Third-party arithmetic used by the reference and absent from /root/reference: PyTorch ATen (pinned 1.11 in
environment.yml:12; here torch 2.11) for conv2d / conv_transpose2d (conv2d_gradfix.py:113-115) -- those two are
called here as well; grid_sample, sort, searchsorted, cumsum, cumprod, softplus are RESTATED below with explicit
index arithmetic, and checked against ATen through the reference run.

Every function cites the reference file:line it follows (paths relative to /root/reference/src).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ======================================================================================================
# torch_utils/ops


def upfirdn2d_out_size(in_size, up, down, pad0, pad1, fsize):
    """Integer output extent, torch_utils/ops/upfirdn2d.cpp:35-36 (C division truncates toward zero)."""
    num = in_size * up + pad0 + pad1 - fsize + down
    return int(math.trunc(num / down))


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    if len(padding) == 2:
        padding = [padding[0], padding[0], padding[1], padding[1]]
    return [int(p) for p in padding]


def _parse_scaling(s):
    return (s, s) if isinstance(s, int) else (int(s[0]), int(s[1]))


FAST_FIR = False   # bench.py's CPU arm sets this: FIR through ATen's grouped conv2d exactly as the reference's own CPU path does


DIFFERENTIABLE = False   # bench.py's EXECUTED CPU training step (oracle/train_step.py) sets this: every op stays a torch op, so autograd can
                         # differentiate the restatement exactly as it differentiates the reference's own CPU path


def _upfirdn2d_aten(x, f, up, down, padding, flip_filter, gain):
    """numpy-in / numpy-out wrapper of `upfirdn2d_t` (CPU timing with FAST_FIR)."""
    ft = None if f is None else torch.as_tensor(np.ascontiguousarray(f), dtype=torch.float32)
    return upfirdn2d_t(torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32), ft, up, down, padding, flip_filter, gain).numpy()


def upfirdn2d_t(xt, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    """torch_utils/ops/upfirdn2d.py:167-211 evaluated the way the reference does on CPU: zero-stuff, pad, grouped F.conv2d, slice.
    Multi-threaded and differentiable; tests/test_cpu_oracle.py checks it against the explicit restatement."""
    N, C, H, W = xt.shape
    upx, upy = _parse_scaling(up); downx, downy = _parse_scaling(down)
    px0, px1, py0, py1 = _parse_padding(padding)
    ft = torch.ones([1, 1]) if f is None else (f if isinstance(f, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(f), dtype=torch.float32))
    xt = xt.reshape([N, C, H, 1, W, 1])
    xt = F.pad(xt, [0, upx - 1, 0, 0, 0, upy - 1]).reshape([N, C, H * upy, W * upx])
    xt = F.pad(xt, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    xt = xt[:, :, max(-py0, 0): xt.shape[2] - max(-py1, 0), max(-px0, 0): xt.shape[3] - max(-px1, 0)]
    ft = ft * (gain ** (ft.ndim / 2))
    if not flip_filter:
        ft = ft.flip(list(range(ft.ndim)))
    ft = ft[None, None].repeat([C, 1] + [1] * ft.ndim)
    if ft.ndim == 4:
        xt = F.conv2d(xt, ft, groups=C)
    else:
        xt = F.conv2d(xt, ft.unsqueeze(2), groups=C)
        xt = F.conv2d(xt, ft.unsqueeze(3), groups=C)
    return xt[:, :, ::downy, ::downx]


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    """torch_utils/ops/upfirdn2d.py:167-211 (`_upfirdn2d_ref`) restated with explicit indexing in numpy float32:
    zero-insert upsample -> pad / crop -> FIR (true convolution unless flip_filter) -> decimate.
    Accumulation order: taps (ky, kx) ascending, float32, gain applied to the filter like the reference."""
    if FAST_FIR:
        return _upfirdn2d_aten(x, f, up, down, padding, flip_filter, gain)
    x = np.asarray(x, dtype=np.float32)
    assert x.ndim == 4
    N, C, H, W = x.shape
    if f is None:
        f = np.ones([1, 1], np.float32)
    f = np.asarray(f, dtype=np.float32)
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    px0, px1, py0, py1 = _parse_padding(padding)
    if f.ndim == 1:   # separable: horizontal pass then vertical pass (upfirdn2d.py:205-207, 243-245)
        g = np.float32(gain) ** np.float32(0.5)
        y = upfirdn2d(x, f[None, :] * 1.0, (upx, 1), (downx, 1), [px0, px1, 0, 0], flip_filter, float(g))
        return upfirdn2d(y, f[:, None] * 1.0, (1, upy), (1, downy), [0, 0, py0, py1], flip_filter, float(g))
    fh, fw = f.shape
    # upsample by zero insertion
    u = np.zeros([N, C, H * upy, W * upx], np.float32)
    u[:, :, ::upy, ::upx] = x
    # pad / crop
    u = np.pad(u, [(0, 0), (0, 0), (max(py0, 0), max(py1, 0)), (max(px0, 0), max(px1, 0))])
    u = u[:, :, max(-py0, 0): u.shape[2] - max(-py1, 0), max(-px0, 0): u.shape[3] - max(-px1, 0)]
    fk = (f * np.float32(gain)).astype(np.float32)
    if not flip_filter:
        fk = fk[::-1, ::-1]
    oh, ow = u.shape[2] - fh + 1, u.shape[3] - fw + 1
    assert oh >= 1 and ow >= 1
    y = np.zeros([N, C, oh, ow], np.float32)
    for ky in range(fh):
        for kx in range(fw):
            y += u[:, :, ky:ky + oh, kx:kx + ow] * fk[ky, kx]
    return y[:, :, ::downy, ::downx]


def setup_filter(f, normalize=True, flip_filter=False, gain=1, separable=None):
    """torch_utils/ops/upfirdn2d.py:70-114."""
    f = np.asarray(1 if f is None else f, dtype=np.float32)
    if f.ndim == 0:
        f = f[None]
    if separable is None:
        separable = (f.ndim == 1 and f.size >= 8)
    if f.ndim == 1 and not separable:
        f = np.outer(f, f)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f[::-1].copy() if f.ndim == 1 else f[::-1, ::-1].copy()
    f = f * (gain ** (f.ndim / 2))
    return f.astype(np.float32)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1):
    """upfirdn2d.py:313-348."""
    upx, upy = _parse_scaling(up)
    px0, px1, py0, py1 = _parse_padding(padding)
    fh, fw = (f.shape[0], f.shape[-1])
    p = [px0 + (fw + upx - 1) // 2, px1 + (fw - upx) // 2, py0 + (fh + upy - 1) // 2, py1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1):
    """upfirdn2d.py:352-387."""
    dx, dy = _parse_scaling(down)
    px0, px1, py0, py1 = _parse_padding(padding)
    fh, fw = (f.shape[0], f.shape[-1])
    p = [px0 + (fw - dx + 1) // 2, px1 + (fw - dx) // 2, py0 + (fh - dy + 1) // 2, py1 + (fh - dy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain)


def filter2d(x, f, padding=0, flip_filter=False, gain=1):
    """upfirdn2d.py:277-309."""
    px0, px1, py0, py1 = _parse_padding(padding)
    fh, fw = (f.shape[0], f.shape[-1])
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain)


ACT = {  # name: (func, def_alpha, def_gain)  -- torch_utils/ops/bias_act.py:21-31
    'linear': (lambda x, a: x, 0, 1.0),
    'relu': (lambda x, a: np.maximum(x, 0), 0, math.sqrt(2)),
    'lrelu': (lambda x, a: np.where(x > 0, x, x * np.float32(a)), 0.2, math.sqrt(2)),
    'tanh': (lambda x, a: np.tanh(x), 0, 1.0),
    'sigmoid': (lambda x, a: 1 / (1 + np.exp(-x)), 0, 1.0),
    'elu': (lambda x, a: np.where(x > 0, x, np.expm1(np.minimum(x, 0))), 0, 1.0),
    'selu': (lambda x, a: 1.0507009873554804934193349852946 * np.where(x > 0, x, 1.6732632423543772848170429916717 * np.expm1(np.minimum(x, 0))), 0, 1.0),
    'softplus': (lambda x, a: np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20)))), 0, 1.0),
    'swish': (lambda x, a: x / (1 + np.exp(-x)), 0, math.sqrt(2)),
}


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """torch_utils/ops/bias_act.py:91-120 (`_bias_act_ref`), numpy float32."""
    x = np.asarray(x, dtype=np.float32)
    fn, da, dg = ACT[act]
    alpha = float(da if alpha is None else alpha)
    gain = float(dg if gain is None else gain)
    if b is not None:
        b = np.asarray(b, dtype=np.float32)
        x = x + b.reshape([-1 if i == dim else 1 for i in range(x.ndim)])
    x = fn(x, alpha).astype(np.float32)
    if gain != 1:
        x = x * np.float32(gain)
    if clamp is not None and clamp >= 0:
        x = np.clip(x, -np.float32(clamp), np.float32(clamp))
    return x.astype(np.float32)


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=math.sqrt(2), slope=0.2, clamp=None, flip_filter=False):
    """torch_utils/ops/filtered_lrelu.py:121-153 (`_filtered_lrelu_ref`)."""
    px0, px1, py0, py1 = _parse_padding(padding)
    x = bias_act(x, b)
    x = upfirdn2d(x, fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    x = bias_act(x, act='lrelu', alpha=slope, gain=gain, clamp=clamp)
    return upfirdn2d(x, fd, down=down, flip_filter=flip_filter)


def _t(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32) if not isinstance(x, torch.Tensor) else x.float()


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """torch_utils/ops/conv2d_resample.py:46-141.  x, w: torch CPU float32; f: numpy.  Dense convs use ATen
    (conv2d_gradfix.py:113-115), FIR passes use the restated upfirdn2d above."""
    x = _t(x); w = _t(w)
    oc, icg, kh, kw = w.shape
    fw, fh = (1, 1) if f is None else (f.shape[-1], f.shape[0])
    px0, px1, py0, py1 = _parse_padding(padding)
    if up > 1:
        px0 += (fw + up - 1) // 2; px1 += (fw - up) // 2; py0 += (fh + up - 1) // 2; py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2; px1 += (fw - down) // 2; py0 += (fh - down + 1) // 2; py1 += (fh - down) // 2

    def conv(x, w, stride=1, padding=0, transpose=False, flip_weight=True):   # conv2d_resample.py:29-41
        if not flip_weight and (w.shape[2] > 1 or w.shape[3] > 1):
            w = w.flip([2, 3])
        if transpose:
            return F.conv_transpose2d(x, w, stride=stride, padding=padding, groups=groups)
        return F.conv2d(x, w, stride=stride, padding=padding, groups=groups)

    def fir(x, **kw_):
        if DIFFERENTIABLE:
            return upfirdn2d_t(x, f, **kw_)
        return _t(upfirdn2d(x.numpy(), f, **kw_))

    if kw == 1 and kh == 1 and down > 1 and up == 1:                         # :94-97
        x = fir(x, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return conv(x, w, flip_weight=flip_weight)
    if kw == 1 and kh == 1 and up > 1 and down == 1:                         # :100-103
        x = conv(x, w, flip_weight=flip_weight)
        return fir(x, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:                                                 # :106-109
        x = fir(x, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return conv(x, w, stride=down, flip_weight=flip_weight)
    if up > 1:                                                               # :112-129
        if groups == 1:
            w = w.transpose(0, 1)
        else:
            w = w.reshape(groups, oc // groups, icg, kh, kw).transpose(1, 2).reshape(groups * icg, oc // groups, kh, kw)
        px0 -= kw - 1; px1 -= kw - up; py0 -= kh - 1; py1 -= kh - up
        pxt = max(min(-px0, -px1), 0); pyt = max(min(-py0, -py1), 0)
        x = conv(x, w, stride=up, padding=[pyt, pxt], transpose=True, flip_weight=(not flip_weight))
        x = fir(x, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = fir(x, down=down, flip_filter=flip_filter)
        return x
    if up == 1 and down == 1 and px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:   # :132-134
        return conv(x, w, padding=[py0, px0], flip_weight=flip_weight)
    if DIFFERENTIABLE:
        x = upfirdn2d_t(x, f if up > 1 else None, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    else:
        x = _t(upfirdn2d(x.numpy(), f if up > 1 else None, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter))
    x = conv(x, w, flip_weight=flip_weight)
    if down > 1:
        x = fir(x, down=down, flip_filter=flip_filter)
    return x


# ======================================================================================================
# training/layers.py, networks_stylegan2.py


_ACT_T = {'linear': lambda x, a: x, 'relu': lambda x, a: torch.relu(x), 'lrelu': lambda x, a: F.leaky_relu(x, a), 'tanh': lambda x, a: torch.tanh(x),
          'sigmoid': lambda x, a: torch.sigmoid(x), 'softplus': lambda x, a: F.softplus(x)}


def t_bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    if DIFFERENTIABLE:   # torch_utils/ops/bias_act.py:91-120 in torch ops (what the reference itself runs on CPU tensors)
        _, da, dg = ACT[act]
        alpha = float(da if alpha is None else alpha); gain = float(dg if gain is None else gain)
        if b is not None:
            x = x + b.reshape([-1 if i == dim else 1 for i in range(x.ndim)])
        x = _ACT_T[act](x, alpha)
        if gain != 1:
            x = x * gain
        if clamp is not None and clamp >= 0:
            x = x.clamp(-clamp, clamp)
        return x
    return _t(bias_act(x.detach().numpy(), None if b is None else b.detach().numpy(), dim, act, alpha, gain, clamp))


def fully_connected(x, weight, bias, activation='linear', lr_multiplier=1.0):
    """training/layers.py:42-58."""
    w = weight * (lr_multiplier / math.sqrt(weight.shape[1]))
    b = bias
    if b is not None and lr_multiplier != 1:
        b = b * lr_multiplier
    if activation == 'linear' and b is not None:
        return torch.addmm(b.unsqueeze(0), x, w.t())
    return t_bias_act(x.matmul(w.t()), b, act=activation)


def normalize_2nd_moment(x, dim=1, eps=1e-8):
    """training/layers.py:16-17."""
    return x * (x.square().mean(dim=dim, keepdim=True) + eps).rsqrt()


def mapping_network(sd, prefix, z, c, num_ws, num_layers=2, lr_multiplier=0.01, truncation_psi=1):
    """training/layers.py:127-174 (no camera conditioning: camera_cond=False in the 3dgp config)."""
    x = None
    if z is not None:
        x = normalize_2nd_moment(z.float())
    if (prefix + 'embed.weight') in sd:
        y = normalize_2nd_moment(fully_connected(c.float(), sd[prefix + 'embed.weight'], sd[prefix + 'embed.bias']))
        x = torch.cat([x, y], dim=1) if x is not None else y
    for i in range(num_layers):
        x = fully_connected(x, sd[f'{prefix}fc{i}.weight'], sd[f'{prefix}fc{i}.bias'], activation='lrelu', lr_multiplier=lr_multiplier)
    if num_ws is not None:
        x = x.unsqueeze(1).repeat([1, num_ws, 1])
    if truncation_psi != 1:
        x = sd[prefix + 'w_avg'].lerp(x, truncation_psi)
    return x


def modulated_conv2d(x, weight, styles, noise=None, up=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True):
    """training/networks_stylegan2.py:31-88 (fp32 branch)."""
    B = x.shape[0]
    oc, ic, kh, kw = weight.shape
    w = None; dcoefs = None
    if demodulate or fused_modconv:
        w = weight.unsqueeze(0) * styles.reshape(B, 1, -1, 1, 1)
    if demodulate:
        dcoefs = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
    if demodulate and fused_modconv:
        w = w * dcoefs.reshape(B, -1, 1, 1, 1)
    if not fused_modconv:
        x = x * styles.reshape(B, -1, 1, 1)
        x = conv2d_resample(x, weight, f=resample_filter, up=up, padding=padding, flip_weight=flip_weight)
        if demodulate and noise is not None:
            x = torch.addcmul(noise, x, dcoefs.reshape(B, -1, 1, 1))     # fma.fma (:71)
        elif demodulate:
            x = x * dcoefs.reshape(B, -1, 1, 1)
        elif noise is not None:
            x = x + noise
        return x
    x = x.reshape(1, -1, *x.shape[2:])
    w = w.reshape(-1, ic, kh, kw)
    x = conv2d_resample(x, w, f=resample_filter, up=up, padding=padding, groups=B, flip_weight=flip_weight)
    x = x.reshape(B, -1, *x.shape[2:])
    if noise is not None:
        x = x + noise
    return x


def synthesis_layer(sd, prefix, x, w, up=1, noise_mode='const', noise_in=None, fused_modconv=True, gain=1, conv_clamp=None, use_noise=True):
    """training/networks_stylegan2.py:128-145."""
    styles = fully_connected(w, sd[prefix + 'affine.weight'], sd[prefix + 'affine.bias'])
    noise = None
    if use_noise and noise_mode == 'random':
        noise = noise_in * sd[prefix + 'noise_strength']
    if use_noise and noise_mode == 'const':
        noise = sd[prefix + 'noise_const'] * sd[prefix + 'noise_strength']
    weight = sd[prefix + 'weight']
    x = modulated_conv2d(x, weight, styles, noise=noise, up=up, padding=weight.shape[-1] // 2,
                         resample_filter=sd[prefix + 'resample_filter'].numpy(), flip_weight=(up == 1), fused_modconv=fused_modconv)
    act_gain = math.sqrt(2) * gain
    return t_bias_act(x, sd[prefix + 'bias'], act='lrelu', gain=act_gain, clamp=(conv_clamp * gain if conv_clamp is not None else None))


def torgb_layer(sd, prefix, x, w, fused_modconv=True, conv_clamp=None):
    """training/networks_stylegan2.py:168-172."""
    weight = sd[prefix + 'weight']
    styles = fully_connected(w, sd[prefix + 'affine.weight'], sd[prefix + 'affine.bias']) * (1 / math.sqrt(weight.shape[1] * weight.shape[2] ** 2))
    x = modulated_conv2d(x, weight, styles, demodulate=False, fused_modconv=fused_modconv)
    return t_bias_act(x, sd[prefix + 'bias'], clamp=conv_clamp)


def tri_plane_decoder(sd, prefix, ws, block_resolutions, noise_mode='const', noises=None, fused_modconv=True):
    """training/networks_epigraf.py:114-129 + networks_stylegan2.py:231-273 ('skip' architecture, fp32)."""
    noises = list(noises) if noises is not None else None
    nxt = (lambda: noises.pop(0)) if noises is not None else (lambda: None)
    x = img = None
    w_idx = 0
    for res in block_resolutions:
        bp = f'{prefix}b{res}.'
        first = (bp + 'const') in sd
        num_conv = 1 if first else 2
        cur = ws[:, w_idx:w_idx + num_conv + 1]
        w_idx += num_conv
        wi = 0
        if first:
            x = sd[bp + 'const'].unsqueeze(0).repeat([ws.shape[0], 1, 1, 1])
            x = synthesis_layer(sd, bp + 'conv1.', x, cur[:, wi], noise_mode=noise_mode, noise_in=nxt() if noise_mode == 'random' else None, fused_modconv=fused_modconv); wi += 1
        else:
            x = synthesis_layer(sd, bp + 'conv0.', x, cur[:, wi], up=2, noise_mode=noise_mode, noise_in=nxt() if noise_mode == 'random' else None, fused_modconv=fused_modconv); wi += 1
            x = synthesis_layer(sd, bp + 'conv1.', x, cur[:, wi], noise_mode=noise_mode, noise_in=nxt() if noise_mode == 'random' else None, fused_modconv=fused_modconv); wi += 1
        if img is not None:
            if DIFFERENTIABLE:     # upsample2d (upfirdn2d.py:313-348) for the 4-tap filter: up 2, pad [2, 1, 2, 1], gain 4
                fl = sd[bp + 'resample_filter']
                p0, p1 = (fl.shape[-1] + 1) // 2, (fl.shape[-1] - 2) // 2
                img = upfirdn2d_t(img, fl, up=2, padding=[p0, p1, p0, p1], gain=4)
            else:
                img = _t(upsample2d(img.numpy(), sd[bp + 'resample_filter'].numpy()))
        y = torgb_layer(sd, bp + 'torgb.', x, cur[:, wi], fused_modconv=fused_modconv)
        img = img + y if img is not None else y
    return img


# ======================================================================================================
# training/rendering_utils.py, tri_plane_renderer.py


def spherical2cartesian(rotation, pitch, radius):
    """training/rendering_utils.py:270-285."""
    x = radius * torch.sin(pitch) * torch.sin(-rotation)
    y = radius * torch.cos(pitch)
    z = radius * torch.sin(pitch) * torch.cos(rotation)
    return torch.stack([x, y, z], dim=-1)


def _normalize(x, dim=-1):
    return x / torch.norm(x, dim=dim, keepdim=True)


def compute_cam2world_matrix(angles, radius, look_at):
    """training/rendering_utils.py:194-218."""
    origins = spherical2cartesian(angles[:, 0], angles[:, 1], radius)
    la = spherical2cartesian(look_at[:, 0], look_at[:, 1], look_at[:, 2])
    fwd = _normalize(_normalize(la - origins))
    up = torch.tensor([0, 1, 0], dtype=torch.float).expand_as(fwd)
    left = _normalize(torch.cross(up, fwd, dim=-1))
    up = _normalize(torch.cross(fwd, left, dim=-1))
    B = fwd.shape[0]
    rot = torch.eye(4).unsqueeze(0).repeat(B, 1, 1)
    rot[:, :3, :3] = torch.stack((-left, up, -fwd), dim=-1)
    tr = torch.eye(4).unsqueeze(0).repeat(B, 1, 1)
    tr[:, :3, 3] = origins
    return tr @ rot


def sample_rays(c2w, fov, resolution, patch_scales=None, patch_offsets=None):
    """training/tri_plane_renderer.py:487-527 (fov: [B] tensor in degrees)."""
    B = len(c2w)
    w, h = resolution
    x, y = torch.meshgrid(torch.linspace(-1, 1, w), torch.linspace(1, -1, h), indexing='ij')
    x = x.T.flatten().unsqueeze(0).repeat(B, 1)
    y = y.T.flatten().unsqueeze(0).repeat(B, 1)
    if patch_scales is not None:
        x = (x + 1.0) * patch_scales[:, 0].view(B, 1) - 1.0 + patch_offsets[:, 0].view(B, 1) * 2.0
        y = (y + 1.0) * patch_scales[:, 1].view(B, 1) - 1.0 + patch_offsets[:, 1].view(B, 1) * 2.0
    fov_rad = fov.unsqueeze(1).expand(B, 1) / 360 * 2 * np.pi
    z = -torch.ones((B, h * w)) / torch.tan(fov_rad * 0.5)
    d_cam = _normalize(torch.stack([x, y, z], dim=2), dim=2)
    d_world = torch.bmm(c2w[..., :3, :3], d_cam.permute(0, 2, 1)).permute(0, 2, 1).reshape(B, h * w, 3)
    ho = torch.zeros((B, 4, h * w)); ho[:, 3, :] = 1
    o_world = torch.bmm(c2w, ho).permute(0, 2, 1).reshape(B, h * w, 4)[..., :3]
    return o_world, d_world


def bilinear_planes(planes, coords):
    """F.grid_sample(bilinear, align_corners=True, zeros) restated (training/tri_plane_renderer.py:575-585).
    planes: [B, 3, C, P, P] ; coords: [B, M, 3] already divided by the box half-size.  Returns [B, 3, M, C]."""
    B, _, C, P, _ = planes.shape
    out = []
    pairs = [(0, 1), (0, 2), (1, 2)]     # (width-coordinate, height-coordinate) of planes xy, xz, yz
    M = coords.shape[1]
    for k, (iu, iv) in enumerate(pairs):
        ix = (coords[..., iu] + 1) / 2 * (P - 1)
        iy = (coords[..., iv] + 1) / 2 * (P - 1)
        x0 = torch.floor(ix); y0 = torch.floor(iy)
        acc = torch.zeros(B, M, C)
        texels = planes[:, k].permute(0, 2, 3, 1).reshape(B, P * P, C)        # row = texel, col = channel
        for dy in (0, 1):
            for dx in (0, 1):
                xx = x0 + dx; yy = y0 + dy
                wx = (x0 + 1 - ix) if dx == 0 else (ix - x0)
                wy = (y0 + 1 - iy) if dy == 0 else (iy - y0)
                valid = (xx >= 0) & (xx <= P - 1) & (yy >= 0) & (yy <= P - 1)
                xi = xx.clamp(0, P - 1).long(); yi = yy.clamp(0, P - 1).long()
                idx = yi * P + xi                                              # [B, M]
                v = torch.stack([texels[b].index_select(0, idx[b]) for b in range(B)])   # [B, M, C]
                acc = acc + v * (wx * wy * valid.float()).unsqueeze(-1)
        out.append(acc)
    return torch.stack(out, dim=1)


def tri_plane_mlp(feats, w1, b1, w2, b2):
    """training/networks_epigraf.py:46-68 (classical marcher: raw rgb)."""
    B, _, M, C = feats.shape
    x = feats.mean(dim=1).reshape(B * M, C)
    x = fully_connected(x, w1, b1, activation='lrelu')
    x = fully_connected(x, w2, b2, activation='linear')
    x = x.view(B, M, -1)
    return x[..., :-1], x[..., [-1]]


def linspace01(N):
    """torch.linspace(0, 1, N) -- ATen's symmetric evaluation (start + i*step for the first half, end - (N-1-i)*step after)."""
    step = np.float32(1.0) / np.float32(N - 1)
    i = np.arange(N)
    lo = (step * i.astype(np.float32)).astype(np.float32)
    hi = (np.float32(1.0) - step * (N - 1 - i).astype(np.float32)).astype(np.float32)
    return torch.from_numpy(np.where(i < N // 2, lo, hi).astype(np.float32))


def sample_stratified(u, N):
    """training/tri_plane_renderer.py:224-230 (classical).  u: [B, R, N] in [0,1)."""
    g = linspace01(N).reshape(1, 1, N)
    mids = 0.5 * (g[..., 1:] + g[..., :-1])
    upper = torch.cat([mids, g[..., -1:]], dim=-1)
    lower = torch.cat([g[..., :1], mids], dim=-1)
    return lower + (upper - lower) * u


def softplus(x):
    """F.softplus(beta=1, threshold=20)."""
    return torch.where(x > 20, x, torch.log1p(torch.exp(torch.clamp(x, max=20))))


def ray_march(colors, densities, depths, use_inf_depth=True, last_back=False, white_back_end_idx=0, clamp_mode='softplus'):
    """training/tri_plane_renderer.py:353-405 with explicit sequential products.  colors [B,R,S,3], densities/depths [B,R,S]."""
    deltas = depths[..., 1:] - depths[..., :-1]
    last = torch.full_like(deltas[..., :1], 1e10 if use_inf_depth else 1e-3)
    deltas = torch.cat([deltas, last], dim=-1)
    sig = softplus(densities) if clamp_mode == 'softplus' else torch.relu(densities)
    alphas = 1.0 - torch.exp(-deltas * sig)
    S = alphas.shape[-1]
    T = torch.ones_like(alphas[..., 0])
    ws = []
    for i in range(S):
        ws.append(alphas[..., i] * T)
        T = T * (1.0 - alphas[..., i] + 1e-10)
    weights = torch.stack(ws, dim=-1)
    wagg = weights.sum(dim=-1)
    if last_back:
        weights = weights.clone(); weights[..., -1] += (1.0 - wagg)
    rgb = (weights.unsqueeze(-1) * colors).sum(dim=-2)
    depth = (weights * depths).sum(dim=-1)
    if white_back_end_idx > 0:
        rgb = rgb.clone(); rgb[..., :white_back_end_idx] = rgb[..., :white_back_end_idx] + 1 - wagg.unsqueeze(-1)
    return rgb, depth, weights, T


def sample_pdf(bins, weights, u, eps=1e-5):
    """training/tri_plane_renderer.py:257-295 with searchsorted(right=True) restated as a count."""
    weights = weights + eps
    pdf = weights / weights.sum(-1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    n_s = weights.shape[1]
    inds = (cdf.unsqueeze(1) <= u.unsqueeze(2)).sum(-1)          # [rays, N_importance]
    below = torch.clamp_min(inds - 1, 0)
    above = torch.clamp_max(inds, n_s)
    c0 = torch.gather(cdf, 1, below); c1 = torch.gather(cdf, 1, above)
    b0 = torch.gather(bins, 1, below); b1 = torch.gather(bins, 1, above)
    den = c1 - c0
    den = torch.where(den < eps, torch.ones_like(den), den)
    return b0 + (u - c0) / den * (b1 - b0)


def render(planes, w1, b1, w2, b2, ray_o, ray_d, u_coarse, u_fine, ray_start, ray_end, box_half, N,
           sn_coarse=None, sn_fine=None, noise_std=0.0, use_inf_depth=True, last_back=False, white_back_end_idx=0,
           clamp_mode='softplus', return_aux=False):
    """ImportanceRenderer.forward -- training/tri_plane_renderer.py:126-170.
    planes [B,3,C,P,P]; ray_o/ray_d [B,R,3]; u_coarse/u_fine [B,R,N].  Returns rgb [B,R,3], depth [B,R], wsum [B,R], T [B,R]."""
    B, R, _ = ray_o.shape
    s2t = lambda s: s * ray_end + (1 - s) * ray_start

    def run(tdist, sn):
        pts = (ray_o.unsqueeze(-2) + tdist.unsqueeze(-1) * ray_d.unsqueeze(-2)).reshape(B, -1, 3) / box_half
        rgb, sigma = tri_plane_mlp(bilinear_planes(planes, pts), w1, b1, w2, b2)
        rgb = rgb.reshape(B, R, N, 3); sigma = sigma.reshape(B, R, N)
        if noise_std > 0:
            sigma = sigma + sn * noise_std
        return rgb, sigma

    s_co = sample_stratified(u_coarse, N)
    t_co = s2t(s_co)
    c_co, d_co = run(t_co, sn_coarse)
    _, _, w_co, _ = ray_march(c_co, d_co, s_co, use_inf_depth, last_back, white_back_end_idx, clamp_mode)   # s-space (:152)
    wts = w_co.detach().reshape(B * R, N) + 1e-5          # importance sampling runs under no_grad + detach (:241, :254)
    z = s_co.reshape(B * R, N)
    zmid = 0.5 * (z[:, :-1] + z[:, 1:])
    s_fi = sample_pdf(zmid, wts[:, 1:-1], u_fine.reshape(B * R, N)).reshape(B, R, N)
    t_fi = s2t(s_fi)
    c_fi, d_fi = run(t_fi, sn_fine)
    all_t = torch.cat([t_co, t_fi], dim=-1)
    all_c = torch.cat([c_co, c_fi], dim=-2)
    all_d = torch.cat([d_co, d_fi], dim=-1)
    order = np.argsort(all_t.detach().numpy(), axis=-1, kind='stable')
    order = torch.from_numpy(order)
    all_t = torch.gather(all_t, -1, order)
    all_d = torch.gather(all_d, -1, order)
    all_c = torch.gather(all_c, -2, order.unsqueeze(-1).expand(-1, -1, -1, 3))
    rgb, depth, weights, T = ray_march(all_c, all_d, all_t, use_inf_depth, last_back, white_back_end_idx, clamp_mode)
    if return_aux:
        return rgb, depth, weights.sum(-1), T, dict(s_coarse=s_co, s_fine=s_fi, sigma_coarse=d_co)
    return rgb, depth, weights.sum(-1), T


# ======================================================================================================
# training/networks_depth_adaptor.py, networks_discriminator.py


def conv2d_layer(sd, prefix, x, activation='linear', up=1, down=1, gain=1, conv_clamp=None, c=None):
    """training/layers.py:228-241."""
    weight = sd[prefix + 'weight']
    w = weight * (1 / math.sqrt(weight.shape[1] * weight.shape[2] ** 2))
    if (prefix + 'affine.weight') in sd:
        mod = 1.0 + fully_connected(c, sd[prefix + 'affine.weight'], sd[prefix + 'affine.bias']).tanh().unsqueeze(2).unsqueeze(3)
        x = x * mod
    x = conv2d_resample(x, w, f=sd[prefix + 'resample_filter'].numpy(), up=up, down=down, padding=weight.shape[-1] // 2, flip_weight=(up == 1))
    act_gain = ACT[activation][2] * gain
    b = sd.get(prefix + 'bias', None)
    return t_bias_act(x, b, act=activation, gain=act_gain, clamp=(conv_clamp * gain if conv_clamp is not None else None))


def depth_adaptor(sd, prefix, depth, cfg, min_depth, max_depth, head_idx=None):
    """training/networks_depth_adaptor.py:49-99; eval mode (last head) unless head_idx [B] is given."""
    B = depth.shape[0]
    raw = sd[prefix + 'near_plane_offset_raw'].repeat(B)
    near = min_depth + raw.sigmoid() * cfg['near_plane_offset_max_fraction'] * (max_depth - min_depth)
    near = near.view(B, 1, 1, 1)
    x = (depth - 0.5 * (max_depth + near)) / ((max_depth - near) + 1e-12) * 2.0
    outs = [x]
    for i in range(cfg['num_hid_layers']):
        x = conv2d_layer(sd, f'{prefix}layers.{i}.', x, activation='lrelu')
        outs.append(conv2d_layer(sd, prefix + 'head.', x, activation='linear'))
    outs = torch.stack(outs).transpose(0, 1)
    if head_idx is None:
        head_idx = torch.full([B], outs.shape[1] - 1, dtype=torch.int64)
    return outs[torch.arange(B), head_idx] + 0.0 * outs.max()


def fourier_scalar_encoder(sd, prefix, x, x_multiplier):
    """training/layers.py:280-299, 327-335 (use_raw=False)."""
    x = x.float() * x_multiplier
    coefs = sd[prefix + 'fourier_encoder.fourier_coefs']
    raw = coefs.view(1, 1, -1) * x.unsqueeze(2)
    out = torch.cat([raw.sin(), raw.cos()], dim=2)
    emb = sd[prefix + 'const_embed.weight'][x.round().long()]
    out = torch.cat([out, emb], dim=2)
    return out.view(x.shape[0], -1)


def minibatch_std(x, group_size=4, num_channels=1):
    """training/networks_discriminator.py:104-120."""
    N, C, H, W = x.shape
    G = min(group_size, N)
    Fc = num_channels
    c = C // Fc
    y = x.reshape(G, -1, Fc, c, H, W)
    y = y - y.mean(dim=0)
    y = y.square().mean(dim=0)
    y = (y + 1e-8).sqrt()
    y = y.mean(dim=[2, 3, 4]).reshape(-1, Fc, 1, 1).repeat(G, 1, H, W)
    return torch.cat([x, y], dim=1)


def discriminator(sd, img, c, patch_scales, patch_offsets, block_resolutions, num_additional_start_blocks, predict_feat=False):
    """training/networks_discriminator.py:256-289, 67-90, 157-181 (fp32)."""
    ppc = torch.cat([patch_scales[:, [0]], patch_offsets], dim=1)
    enc = fourier_scalar_encoder(sd, 'scalar_enc.', ppc, 1000.0)
    cc = torch.cat([c, enc], dim=1)
    hyper_c = mapping_network(sd, 'hyper_mod_mapping.', None, enc, None)
    x = None
    s = math.sqrt(0.5)
    for i, res in enumerate(block_resolutions):
        bp = f'b{res}.'
        down = 1 if i < num_additional_start_blocks else 2
        if i == 0:
            x = conv2d_layer(sd, bp + 'fromrgb.', img, activation='lrelu')
        y = conv2d_layer(sd, bp + 'skip.', x, down=down, gain=s)
        x = conv2d_layer(sd, bp + 'conv0.', x, activation='lrelu')
        x = conv2d_layer(sd, bp + 'conv1.', x, activation='lrelu', down=down, gain=s, c=hyper_c)
        x = y + x
    cmap = mapping_network(sd, 'head_mapping.', None, cc, None)
    x = minibatch_std(x)
    x = conv2d_layer(sd, 'b4.conv.', x, activation='lrelu')
    x = x.flatten(1)
    f = None
    if predict_feat:
        f = fully_connected(x, sd['b4.feat_out.0.weight'], sd['b4.feat_out.0.bias'], activation='lrelu')
        f = fully_connected(f, sd['b4.feat_out.1.weight'], sd['b4.feat_out.1.bias'])
    x = fully_connected(x, sd['b4.fc.weight'], sd['b4.fc.bias'], activation='lrelu')
    x = fully_connected(x, sd['b4.out.weight'], sd['b4.out.bias'])
    x = (x * cmap).sum(dim=1, keepdim=True) * (1 / math.sqrt(cmap.shape[1]))
    return x.squeeze(1), f


def generator_synthesis(sd, cfg, ws, angles, fov, radius, look_at, res, patch_scales=None, patch_offsets=None,
                        u_coarse=None, u_fine=None, noise_mode='const', noises=None, fused_modconv=True,
                        depth_head_idx=None, concat_depth=True):
    """training/networks_epigraf.py:210-261 (SynthesisNetwork.forward) for given injected renderer variates."""
    tri = cfg['tri_plane']
    block_res = [2 ** i for i in range(2, int(math.log2(tri['res'])) + 1)]
    dec = tri_plane_decoder(sd, 'synthesis.tri_plane_decoder.', ws, block_res, noise_mode, noises, fused_modconv)
    B = ws.shape[0]
    planes = dec[:, :3 * tri['feat_dim']].view(B, 3, tri['feat_dim'], tri['res'], tri['res'])
    c2w = compute_cam2world_matrix(angles, radius, look_at)
    ro, rd = sample_rays(c2w, fov, (res, res), patch_scales, patch_offsets)
    mp = 'synthesis.tri_plane_mlp.model.'
    rgb, depth, _, _ = render(planes, sd[mp + '0.weight'], sd[mp + '0.bias'], sd[mp + '1.weight'], sd[mp + '1.bias'],
                              ro, rd, u_coarse, u_fine, cfg['camera']['ray']['start'], cfg['camera']['ray']['end'],
                              cfg['camera']['cube_scale'], cfg['num_ray_steps'], use_inf_depth=cfg['use_inf_depth'],
                              last_back=cfg['dataset']['last_back'])
    img = rgb.reshape(B, res, res, 3).permute(0, 3, 1, 2).contiguous()
    dep = depth.reshape(B, 1, res, res)
    out = dict(planes=dec, img=img, depth=dep)
    if cfg['depth_adaptor']['enabled']:
        da = depth_adaptor(sd, 'synthesis.depth_adaptor.', dep, cfg['depth_adaptor'], cfg['camera']['ray']['start'],
                           cfg['camera']['ray']['end'], depth_head_idx)
        out['depth_adapted'] = da
        out['img'] = torch.cat([img, da], dim=1) if concat_depth else img + 0.0 * da.max()
    return out
