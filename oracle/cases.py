"""TEST INFRASTRUCTURE ONLY -- seeded, machine-independent parity cases (numpy RandomState; no torch RNG).
Used by oracle/make_golden.py (to run the reference) and by tests/ (to feed identical inputs to the CUDA path)."""
import zlib

import numpy as np


def _rs(name, salt=0):
    return np.random.RandomState((zlib.crc32(name.encode()) + salt) % (2 ** 31))


def cotangent(shape, seed):
    n = int(np.prod(shape))
    return np.cos(np.arange(n, dtype=np.float64) * 0.37 + seed).astype(np.float32).reshape(shape)


# ----------------------------------------------------------------------------------------------
F1331 = [1, 3, 3, 1]


def _f2d(taps, gain=1.0):
    f = np.asarray(taps, np.float32)
    f = np.outer(f, f)
    return (f / f.sum() * gain).astype(np.float32)


def upfirdn2d_cases():
    """(name, kwargs).  Shapes that occur in G/D + edge cases: negative padding (crop), odd sizes,
    asymmetric factors, 1x1 filter, separable long filter, flip, single pixel, integer-valued (bit-exact) inputs."""
    c = []
    c.append(('g_conv_up2', dict(shape=(2, 5, 17, 17), f='2d4', up=1, down=1, padding=[1, 1, 1, 1], flip_filter=False, gain=4)))      # conv-path FIR after transpose conv
    c.append(('g_skip_up2', dict(shape=(2, 6, 8, 8), f='2d4', up=2, down=1, padding=[2, 1, 2, 1], flip_filter=False, gain=4)))        # upsample2d
    c.append(('d_down2', dict(shape=(2, 4, 16, 16), f='2d4', up=1, down=2, padding=[1, 1, 1, 1], flip_filter=False, gain=1)))        # 1x1 skip downsample
    c.append(('d_pad2', dict(shape=(1, 3, 16, 16), f='2d4', up=1, down=1, padding=[2, 2, 2, 2], flip_filter=False, gain=1)))         # FIR before stride-2 conv
    c.append(('crop', dict(shape=(1, 2, 12, 11), f='2d4', up=1, down=1, padding=[-1, 2, 0, -2], flip_filter=False, gain=1)))
    c.append(('odd_up3_down2', dict(shape=(1, 2, 7, 9), f='2d5', up=3, down=2, padding=[2, 3, 1, 4], flip_filter=False, gain=2.5)))
    c.append(('asym', dict(shape=(1, 3, 6, 10), f='2d3x5', up=[2, 1], down=[1, 3], padding=[1, 2, 3, 0], flip_filter=True, gain=1)))
    c.append(('identity', dict(shape=(1, 1, 5, 5), f=None, up=1, down=1, padding=0, flip_filter=False, gain=1)))
    c.append(('blur_sep61', dict(shape=(1, 4, 64, 64), f='sep61', up=1, down=1, padding=[30, 30, 30, 30], flip_filter=False, gain=1)))  # loss.py:331-337 blur
    c.append(('one_pixel', dict(shape=(1, 1, 1, 1), f='2d4', up=2, down=1, padding=[2, 1, 2, 1], flip_filter=False, gain=4)))
    c.append(('int_up2', dict(shape=(1, 3, 9, 9), f='2d4', up=2, down=1, padding=[2, 1, 2, 1], flip_filter=False, gain=4, integer=True)))
    c.append(('int_down2', dict(shape=(2, 2, 16, 12), f='2d4', up=1, down=2, padding=[1, 1, 1, 1], flip_filter=False, gain=1, integer=True)))
    c.append(('int_flip', dict(shape=(1, 2, 8, 8), f='2d3x5i', up=2, down=2, padding=[3, 1, 0, 2], flip_filter=True, gain=2, integer=True)))
    c.append(('wide', dict(shape=(1, 2, 5, 300), f='2d4', up=2, down=1, padding=[2, 1, 2, 1], flip_filter=False, gain=4)))
    return c


def upfirdn2d_inputs(name, kw):
    rs = _rs('upfirdn2d/' + name)
    if kw.get('integer', False):
        x = rs.randint(-8, 9, size=kw['shape']).astype(np.float32)
    else:
        x = rs.standard_normal(kw['shape']).astype(np.float32)
    fk = kw['f']
    if fk is None:
        f = None
    elif fk == '2d4':
        f = _f2d(F1331)                      # [1,3,3,1] x [1,3,3,1] / 64: dyadic, exact in fp32
    elif fk == '2d5':
        f = _f2d([1, 4, 6, 4, 1])
    elif fk == '2d3x5':
        f = rs.standard_normal((3, 5)).astype(np.float32)
    elif fk == '2d3x5i':
        f = (rs.randint(-3, 4, size=(3, 5)) / 8.0).astype(np.float32)
    elif fk == 'sep61':
        t = np.arange(-30, 31, dtype=np.float32) / 10.0
        f = np.exp2(-t * t).astype(np.float32)
        f = (f / f.sum()).astype(np.float32)
    else:
        raise KeyError(fk)
    return x, f


# ----------------------------------------------------------------------------------------------
def bias_act_cases():
    c = []
    acts = ['linear', 'relu', 'lrelu', 'tanh', 'sigmoid', 'elu', 'selu', 'softplus', 'swish']
    for a in acts:
        c.append((f'{a}_nchw', dict(shape=(2, 6, 5, 7), dim=1, act=a, bias=True, clamp=None)))
    c.append(('lrelu_clamp', dict(shape=(2, 8, 4, 4), dim=1, act='lrelu', bias=True, clamp=0.6, gain=1.2, alpha=0.1)))
    c.append(('lrelu_2d', dict(shape=(9, 64), dim=1, act='lrelu', bias=True, clamp=None)))              # FC / MLP usage
    c.append(('linear_nobias_gain', dict(shape=(3, 5, 2, 2), dim=1, act='linear', bias=False, clamp=None, gain=0.5)))
    c.append(('sigmoid_dim0', dict(shape=(4, 3), dim=0, act='sigmoid', bias=True, clamp=0.7)))
    c.append(('swish_clamp', dict(shape=(2, 4, 3, 3), dim=1, act='swish', bias=True, clamp=1.0)))
    c.append(('tail_odd', dict(shape=(1, 3, 7, 5), dim=1, act='lrelu', bias=True, clamp=None)))         # numel % 4 != 0
    return c


def bias_act_inputs(name, kw):
    rs = _rs('bias_act/' + name)
    x = (rs.standard_normal(kw['shape']) * 1.5).astype(np.float32)
    b = (rs.standard_normal(kw['shape'][kw['dim']]) * 0.5).astype(np.float32) if kw['bias'] else None
    return x, b


# ----------------------------------------------------------------------------------------------
def filtered_lrelu_cases():
    return [
        ('up2_down2', dict(shape=(2, 3, 8, 8), up=2, down=2, fu='k12', fd='k12', padding=[11, 10, 11, 10], gain=np.sqrt(2), slope=0.2, clamp=None)),
        ('up2_down1_clamp', dict(shape=(1, 4, 6, 10), up=2, down=1, fu='2d4', fd=None, padding=[2, 1, 2, 1], gain=1.3, slope=0.1, clamp=0.4)),
        ('up1_down2', dict(shape=(1, 2, 12, 12), up=1, down=2, fu=None, fd='2d4', padding=[1, 1, 1, 1], gain=np.sqrt(2), slope=0.2, clamp=None)),
        # separable filters at sizes that span several 32 x 32 output tiles of the fused kernel (ragged edges, clamp codes, up != down)
        ('sep_40x37_clamp', dict(shape=(2, 5, 40, 37), up=2, down=2, fu='k12', fd='k12', padding=[11, 10, 11, 10], gain=np.sqrt(2), slope=0.2, clamp=0.5)),
        ('sep_up4_down2', dict(shape=(1, 3, 20, 24), up=4, down=2, fu='k12', fd='k6', padding=[9, 8, 9, 8], gain=1.1, slope=0.3, clamp=None)),
        ('sep_up4_down2_k24', dict(shape=(1, 2, 24, 19), up=4, down=2, fu='k24', fd='k12', padding=[21, 20, 21, 20], gain=np.sqrt(2), slope=0.2, clamp=None)),
        ('sep_up2_down1', dict(shape=(1, 3, 30, 41), up=2, down=1, fu='k12', fd=None, padding=[6, 5, 6, 5], gain=np.sqrt(2), slope=0.2, clamp=1.0)),
        ('sep_up1_down1_crop', dict(shape=(1, 2, 70, 33), up=1, down=1, fu='k6', fd='k6', padding=[-1, 2, 3, -2], gain=np.sqrt(2), slope=0.2, clamp=None)),
    ]


def filtered_lrelu_inputs(name, kw):
    rs = _rs('filtered_lrelu/' + name)
    x = rs.standard_normal(kw['shape']).astype(np.float32)
    b = (rs.standard_normal(kw['shape'][1]) * 0.3).astype(np.float32)

    def mk(k):
        if k is None:
            return np.ones([1, 1], np.float32)
        if k == '2d4':
            return _f2d(F1331)
        if k == 'k12':   # 12-tap windowed-sinc-ish separable low-pass (StyleGAN3-like), stored 1-D
            t = np.arange(12, dtype=np.float64) - 5.5
            f = np.sinc(t / 2.0) * np.kaiser(12, 6.0)
            return (f / f.sum()).astype(np.float32)
        if k == 'k24':
            t = np.arange(24, dtype=np.float64) - 11.5
            f = np.sinc(t / 4.0) * np.kaiser(24, 6.0)
            return (f / f.sum()).astype(np.float32)
        if k == 'k6':    # asymmetric 6-tap separable filter (convolution vs correlation matters)
            return (np.array([1, 4, 7, 5, 2, 1], np.float64) / 20.0).astype(np.float32)
        raise KeyError(k)
    return x, mk(kw['fu']), mk(kw['fd']), b


# ----------------------------------------------------------------------------------------------
def render_cases():
    base = dict(C=32, H=64, ray_start=0.75, ray_end=1.25, box_half=0.5)
    return [
        ('base', dict(base, B=2, R=40, N=12, P=32)),
        ('n48', dict(base, B=1, R=33, N=48, P=64)),
        ('noise_lastback', dict(base, B=2, R=17, N=8, P=16, noise_std=0.5, last_back=True)),
        ('white_relu_finite', dict(base, B=1, R=20, N=6, P=16, white_back_end_idx=3, clamp_mode='relu', use_inf_depth=False)),
        ('oob', dict(base, B=1, R=24, N=10, P=16, box_half=0.3)),      # samples leave the cube: zero-padding taps
        ('n3_min', dict(base, B=1, R=5, N=3, P=8)),
    ]


def render_inputs(name, kw):
    rs = _rs('render/' + name)
    B, Rr, N, P, C, H = kw['B'], kw['R'], kw['N'], kw['P'], kw['C'], kw['H']
    planes = rs.standard_normal((B, 3, C, P, P)).astype(np.float32)
    # cameras on the unit sphere looking at the origin-ish, fov ~ U[10,45] deg
    o = rs.standard_normal((B, 1, 3)); o = o / np.linalg.norm(o, axis=-1, keepdims=True)
    tgt = rs.uniform(-0.1, 0.1, size=(B, Rr, 3))
    spread = np.tan(np.deg2rad(rs.uniform(10, 45, size=(B, 1, 1))) / 2)
    d = (tgt - o) + rs.uniform(-1, 1, size=(B, Rr, 3)) * spread
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    out = dict(
        planes=planes, ray_o=np.broadcast_to(o, (B, Rr, 3)).astype(np.float32).copy(), ray_d=d.astype(np.float32),
        w1=rs.standard_normal((H, C)).astype(np.float32), b1=(rs.standard_normal(H) * 0.2).astype(np.float32),
        w2=rs.standard_normal((4, H)).astype(np.float32), b2=(rs.standard_normal(4) * 0.2).astype(np.float32),
        u_coarse=rs.uniform(0, 1, size=(B, Rr, N)).astype(np.float32), u_fine=rs.uniform(0, 1, size=(B, Rr, N)).astype(np.float32))
    if kw.get('noise_std', 0.0) > 0:
        out['sn_coarse'] = rs.standard_normal((B, Rr, N)).astype(np.float32)
        out['sn_fine'] = rs.standard_normal((B, Rr, N)).astype(np.float32)
    return out


# ----------------------------------------------------------------------------------------------
def small_net_kwargs():
    """Reduced-width 3DGP config (same topology as configs/model/3dgp.yaml: skip decoder, 32-ch tri-planes,
    32->64->4 MLP, depth adaptor, hyper-modulated patch D with 2 additional start blocks)."""
    return dict(cmax=32, cbase=512, tri_res=32, feat_dim=32, num_ray_steps=8, patch_res=16, img_resolution=64, c_dim=5,
                use_depth=True, learn_camera_dist=False, hid_dim=64, w_dim=64, z_dim=64, depth_hid=8, embedding_dim=16)


def wide_net_kwargs():
    """Same topology at widths where EVERY hot convolution is eligible for the tcgen05 path (Cin % 64 == 0, Cout in {64, 96} or % 128 == 0):
    G decoder 128-128-128-128-64 channels up to 64^2 tri-planes (incl. a 128->64 up-sampling layer, 64->96 and 128->96 toRGB), depth adaptor at its
    production width (64 ch, 5x5), D 64/128 channels on a 32^2 patch of a 128^2 image: blocks b128..b16 in the reference's fp16 set, b8 / b4 fp32."""
    return dict(cmax=128, cbase=4096, tri_res=64, feat_dim=32, num_ray_steps=12, patch_res=32, img_resolution=128, c_dim=5,
                use_depth=True, learn_camera_dist=False, hid_dim=64, w_dim=64, z_dim=64, depth_hid=64, embedding_dim=16, d_fmaps=2.0)


def net_kwargs(variant):
    return dict(small=small_net_kwargs, wide=wide_net_kwargs)[variant]()


_KEEP_BUFFERS = ('resample_filter', 'fourier_coefs', 'progress_coef')


def fill_state_dict(shapes, ref_sd, seed):
    """Deterministic weights: every parameter ~ N(0,1) (biases/strengths smaller), structural buffers kept."""
    import torch
    rs = np.random.RandomState(seed)
    sd = {}
    for k in sorted(shapes):
        shp = shapes[k]
        if any(k.endswith(s) for s in _KEEP_BUFFERS):
            sd[k] = ref_sd[k].clone()
            continue
        v = rs.standard_normal(shp).astype(np.float32)
        if k.endswith('noise_strength'):
            v = np.asarray(0.1 + 0.05 * v, np.float32)
        elif k.endswith('affine.bias') and 'synthesis' in k:
            v = (1.0 + 0.1 * v).astype(np.float32)
        elif k.endswith('.bias'):
            v = (0.1 * v).astype(np.float32)
        elif k.endswith('near_plane_offset_raw'):
            v = np.asarray([-3.0], np.float32)
        elif k.endswith('w_avg'):
            v = (0.05 * v).astype(np.float32)
        sd[k] = torch.from_numpy(np.ascontiguousarray(v)).reshape(shp)
    return sd


def net_inputs(kw, B=4):
    rs = _rs('net_inputs')
    c = np.zeros((B, kw['c_dim']), np.float32)
    c[np.arange(B), np.arange(B) % kw['c_dim']] = 1
    pr = kw['patch_res']; N = kw['num_ray_steps']
    return dict(
        z=rs.standard_normal((B, kw['z_dim'])).astype(np.float32), c=c,
        angles=np.stack([rs.uniform(-1.2, 1.2, B), rs.uniform(0.9, 2.2, B), np.zeros(B)], 1).astype(np.float32),
        fov=rs.uniform(12, 40, B).astype(np.float32), radius=np.ones(B, np.float32),
        look_at=np.stack([rs.uniform(-3, 3, B), rs.uniform(0.2, 2.9, B), rs.uniform(0, 0.2, B)], 1).astype(np.float32),
        patch_scales=np.full((B, 2), 0.5, np.float32), patch_offsets=np.stack([np.full(B, 0.25), np.full(B, 0.375)], 1).astype(np.float32),
        u_coarse=rs.uniform(0, 1, (B, pr * pr, N)).astype(np.float32), u_fine=rs.uniform(0, 1, (B, pr * pr, N)).astype(np.float32))


def layer_noises(kw, B):
    """Per-layer noise images in the reference's call order: b4.conv1, then conv0, conv1 per block (networks_stylegan2.py:134)."""
    rs = _rs('layer_noises')
    res_list = [2 ** i for i in range(2, int(np.log2(kw['tri_res'])) + 1)]
    out = []
    for r in res_list:
        n = 1 if r == 4 else 2
        for _ in range(n):
            out.append(rs.standard_normal((B, 1, r, r)).astype(np.float32))
    return out


def depth_heads(B):
    return (np.arange(B) % 4).astype(np.int64)


def eval_variates(kw, B):
    rs = _rs('eval_variates')
    Re = kw['img_resolution'] ** 2; N = kw['num_ray_steps']
    return dict(u_coarse=rs.uniform(0, 1, (B, Re, N)).astype(np.float32), u_fine=rs.uniform(0, 1, (B, Re, N)).astype(np.float32))


def grad_probe(g, limit=4096):
    """Strided sample of a (possibly large) gradient tensor: keeps the fixtures small; tests apply the same stride to the CUDA result."""
    g = np.asarray(g).reshape(-1)
    return g[::max(1, g.size // limit)].copy()


def probe_params(which, variant='small'):
    if variant == 'wide':
        if which == 'D':
            return ['b128.fromrgb.weight', 'b128.conv0.weight', 'b128.conv1.weight', 'b128.conv1.affine.weight', 'b128.skip.weight', 'b64.conv0.bias',
                    'b32.conv1.weight', 'b32.skip.weight', 'b16.conv0.weight', 'b8.conv1.weight', 'b8.skip.weight', 'b4.fc.weight', 'b4.out.bias',
                    'head_mapping.fc1.weight', 'hyper_mod_mapping.embed.weight']
        return ['synthesis.tri_plane_mlp.model.0.weight', 'synthesis.tri_plane_mlp.model.1.bias', 'synthesis.tri_plane_decoder.b4.const',
                'synthesis.tri_plane_decoder.b8.conv0.weight', 'synthesis.tri_plane_decoder.b8.conv0.noise_strength',
                'synthesis.tri_plane_decoder.b16.conv1.weight', 'synthesis.tri_plane_decoder.b32.conv1.affine.weight',
                'synthesis.tri_plane_decoder.b32.torgb.weight', 'synthesis.tri_plane_decoder.b32.torgb.affine.bias',
                'synthesis.tri_plane_decoder.b64.conv0.weight', 'synthesis.tri_plane_decoder.b64.conv1.weight', 'synthesis.tri_plane_decoder.b64.conv1.bias',
                'synthesis.tri_plane_decoder.b64.torgb.weight', 'synthesis.depth_adaptor.layers.0.weight', 'synthesis.depth_adaptor.layers.1.weight',
                'synthesis.depth_adaptor.head.weight', 'mapping.fc0.weight']
    if which == 'D':
        return ['b64.fromrgb.weight', 'b64.conv1.affine.weight', 'b16.conv0.weight', 'b8.skip.weight', 'b4.fc.weight', 'b4.out.bias',
                'head_mapping.fc1.weight', 'hyper_mod_mapping.embed.weight']
    return ['synthesis.tri_plane_mlp.model.0.weight', 'synthesis.tri_plane_mlp.model.1.bias', 'synthesis.tri_plane_decoder.b32.conv1.weight',
            'synthesis.tri_plane_decoder.b32.torgb.affine.bias', 'synthesis.tri_plane_decoder.b4.const', 'synthesis.tri_plane_decoder.b8.conv0.noise_strength',
            'synthesis.depth_adaptor.layers.0.weight', 'mapping.fc0.weight']


# ----------------------------------------------------------------------------------------------
def snapshot_fill(name, shape):
    """Low-entropy deterministic tensor content for the snapshot fixture (compresses ~100x): value depends on the parameter name and the flat index."""
    n = int(np.prod(shape)) if len(shape) else 1
    salt = sum(ord(ch) for ch in name) % 11
    return (((np.arange(n, dtype=np.int64) + salt) % 13 - 6).astype(np.float32) / 8.0).reshape(shape)
