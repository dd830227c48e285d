"""Numpy restatement of the tensor-core entry points' CONTRACT as include/gp3d_b200.h words it (gp3d_conv_nhwc / gp3d_conv_taps_nhwc,
gp3d_conv_transpose_s2_nhwc, gp3d_wgrad_taps_nhwc_fmt, gp3d_split_pad), in float64 and without any precision splitting.

Test infrastructure only: tests/test_cpu_tap_algebra.py swaps these in for the ctypes launchers of 3dgp_b200/torch_utils/ops/tc.py so that the HOST
logic above the C ABI -- tap lists, traversal strides, output lattices, weight re-layouts, padding algebra -- runs on a machine without a GPU and is
compared with torch.nn.functional.  Nothing in the product imports this file."""
import numpy as np
import torch


def split(x_nhwc, styles=None, want_lo=True, pad_to=None, fp16=False):
    """gp3d_split_pad: (hi, lo) with hi + lo == x * styles; here hi carries the full value and lo is zero (or absent)."""
    x = x_nhwc.to(torch.float64)
    if styles is not None:
        x = x * styles.to(torch.float64).reshape([x.shape[0]] + [1] * (x.dim() - 2) + [x.shape[-1]])
    C = x.shape[-1]
    if pad_to is not None and int(pad_to) > C:
        x = torch.cat([x, torch.zeros(list(x.shape[:-1]) + [int(pad_to) - C], dtype=x.dtype)], dim=-1)
    x = x.contiguous()
    return x, (torch.zeros_like(x) if want_lo else None)


def _sum(hi, lo):
    """hi (+ lo) as a float64 array: operand pairs are summed in float64 (they may be bf16 tensors written by the emulated backward kernels)."""
    v = hi.to(torch.float64)
    return (v + lo.to(torch.float64) if lo is not None else v).numpy()


def _shifted(x, dy, dx, stride, HoP, WoP):
    """x[n][iy*stride+dy][ix*stride+dx][:] for (iy, ix) in [0,HoP) x [0,WoP), zero outside the tensor -> [N, HoP, WoP, C]."""
    N, H, W, C = x.shape
    out = np.zeros([N, HoP, WoP, C], dtype=np.float64)
    ys = np.arange(HoP) * stride + dy
    xs = np.arange(WoP) * stride + dx
    vy = (ys >= 0) & (ys < H); vx = (xs >= 0) & (xs < W)
    if vy.any() and vx.any():
        out[np.ix_(np.arange(N), np.nonzero(vy)[0], np.nonzero(vx)[0])] = x[np.ix_(np.arange(N), ys[vy], xs[vx])]
    return out


def conv_launch(xh, xl, wh, wl, y, N, H, W, Cin, Cout, slabs, taps, in_stride, HoP, WoP, Hout, Wout, osy, osx, oy0, ox0, epi=None, what=''):
    """y[n][iy*osy+oy0][ix*osx+ox0][co] = sum_t sum_ci x[n][iy*in_stride+dy_t][ix*in_stride+dx_t][ci] * w[co][slab_t][ci]  (gp3d_conv_taps_nhwc)."""
    assert epi is None, 'the fused epilogue is not part of the tap algebra under test'
    x = _sum(xh, xl).reshape(N, H, W, Cin)
    w = _sum(wh, wl).reshape(Cout, slabs, Cin)
    assert tuple(y.shape) == (N, Hout, Wout, Cout)
    acc = np.zeros([N, HoP, WoP, Cout], dtype=np.float64)
    for (dy, dx, slab) in taps:
        assert 0 <= slab < slabs
        acc += _shifted(x, dy, dx, in_stride, HoP, WoP) @ w[:, slab, :].T
    assert oy0 + (HoP - 1) * osy < Hout and ox0 + (WoP - 1) * osx < Wout, 'output lattice leaves the tensor'
    y[:, oy0:oy0 + (HoP - 1) * osy + 1:osy, ox0:ox0 + (WoP - 1) * osx + 1:osx, :] = torch.from_numpy(acc).to(y.dtype)


def conv_transpose_s2_launch(xh, xl, wh, wl, y, N, H, W, Cin, Cout):
    """y[n][2i+ky][2j+kx][co] += x[n][i][j][ci] * w[co][ky*3+kx][ci], y [N][2H+1][2W+1][Cout] fully written  (gp3d_conv_transpose_s2_nhwc)."""
    x = _sum(xh, xl).reshape(N, H, W, Cin)
    w = _sum(wh, wl).reshape(Cout, 9, Cin)
    out = np.zeros([N, 2 * H + 1, 2 * W + 1, Cout], dtype=np.float64)
    for ky in range(3):
        for kx in range(3):
            out[:, ky:ky + 2 * H:2, kx:kx + 2 * W:2, :] += x @ w[:, ky * 3 + kx, :].T
    y.copy_(torch.from_numpy(out).to(y.dtype))


def wgrad_launch(dh, dl, xh, xl, dW, N, Hd, Wd, Cy, Hx, Wx, Cx, slabs, taps, sa, sb, HoP, WoP):
    """dW[co][slab_t][ci] += sum_{n,iy,ix} dy[n][iy*sa+ay_t][ix*sa+ax_t][co] * x[n][iy*sb+by_t][ix*sb+bx_t][ci]  (gp3d_wgrad_taps_nhwc_fmt)."""
    d = _sum(dh, dl).reshape(N, Hd, Wd, Cy)
    x = _sum(xh, xl).reshape(N, Hx, Wx, Cx)
    assert tuple(dW.shape) == (Cy, slabs, Cx)
    acc = dW.to(torch.float64).numpy().copy()
    for (ay, ax, by, bx, slab) in taps:
        a = _shifted(d, ay, ax, sa, HoP, WoP).reshape(-1, Cy)
        b = _shifted(x, by, bx, sb, HoP, WoP).reshape(-1, Cx)
        acc[:, slab, :] += a.T @ b
    dW.copy_(torch.from_numpy(acc).to(dW.dtype))


# ---------------------------------------------------------------------------------------------------------------------------------
# Pointer-level entry points used by the fused layer nodes (3dgp_b200/torch_utils/ops/modconv.py).  CPU tensors hand out host addresses from
# data_ptr(), so the same ctypes call sites can be served by numpy views of those addresses.
import ctypes


def _f32(ptr, *shape):
    n = int(np.prod(shape))
    return np.ctypeslib.as_array(ctypes.cast(int(ptr), ctypes.POINTER(ctypes.c_float)), shape=(n,)).reshape(shape)


def _store_bf16_pair(hi_ptr, lo_ptr, v, C_pad):
    """v [..., C] float64 -> bf16 (hi, lo) with hi + lo ~= v, zero-padded to C_pad channels (gp3d_split_pad / *_bwd_split contract)."""
    lead, C = v.shape[:-1], v.shape[-1]
    full = np.zeros(list(lead) + [C_pad], dtype=np.float32); full[..., :C] = v
    t = torch.from_numpy(full)
    hi = t.to(torch.bfloat16)
    n = full.size
    np.ctypeslib.as_array(ctypes.cast(int(hi_ptr), ctypes.POINTER(ctypes.c_int16)), shape=(n,))[:] = hi.view(torch.int16).numpy().reshape(-1)
    if lo_ptr:
        lo = (t - hi.to(torch.float32)).to(torch.bfloat16)
        np.ctypeslib.as_array(ctypes.cast(int(lo_ptr), ctypes.POINTER(ctypes.c_int16)), shape=(n,))[:] = lo.view(torch.int16).numpy().reshape(-1)


def _act(v, act, alpha, gain, clamp):
    alpha, gain, clamp = float(np.float32(alpha)), float(np.float32(gain)), float(np.float32(clamp))       # `float` parameters of the C ABI
    if act == 3:
        v = np.where(v > 0, v, v * alpha)
    else:
        assert act == 1
    v = v * gain
    return np.clip(v, -clamp, clamp) if clamp > 0 else v


def _epilogue(v, epi, N, H, W, C):
    """act(v * dcoef[n][c] + noise[n?][y][x] + bias[c]) * gain on v [N, H, W, C]  (gp3d_conv_epilogue)."""
    if epi.dcoef:
        v = v * _f32(epi.dcoef, N, 1, 1, C)
    if epi.noise:
        v = v + (_f32(epi.noise, N, H, W, 1) if epi.noise_per_sample else _f32(epi.noise, 1, H, W, 1))
    if epi.bias:
        v = v + _f32(epi.bias, 1, 1, 1, C)
    return _act(v, epi.act, epi.alpha, epi.gain, epi.clamp)


def conv_launch_epi(xh, xl, wh, wl, y, N, H, W, Cin, Cout, slabs, taps, in_stride, HoP, WoP, Hout, Wout, osy, osx, oy0, ox0, epi=None, what=''):
    """conv_launch with the optional fused epilogue; operands may be bf16 pairs written by the emulated backward kernels (summed in float64)."""
    x, w = torch.from_numpy(_sum(xh, xl)), torch.from_numpy(_sum(wh, wl))
    yn = y if y.dim() == 4 and tuple(y.shape) == (N, Hout, Wout, Cout) else y.permute(0, 2, 3, 1)      # _ConvBiasAct passes an NCHW-shaped channels-last tensor
    tmp = torch.zeros([N, Hout, Wout, Cout], dtype=torch.float64)
    conv_launch(x, None, w, None, tmp, N, H, W, Cin, Cout, slabs, taps, in_stride, HoP, WoP, Hout, Wout, osy, osx, oy0, ox0)
    v = tmp.numpy()
    if epi is not None:
        assert (HoP, WoP, osy, osx, oy0, ox0) == (Hout, Wout, 1, 1, 0, 0)
        v = _epilogue(v, epi, N, Hout, Wout, Cout)
    yn.copy_(torch.from_numpy(v).to(y.dtype))


class FakeLib:
    """Stands in for the ctypes CDLL: the entry points ops/modconv.py calls directly, served from host memory per include/gp3d_b200.h."""

    def __getattr__(self, name):            # anything else (bias_act, upfirdn2d, ...) is not part of the node under test
        raise AttributeError(f'abi_emulator.FakeLib: {name} is not emulated')

    @staticmethod
    def gp3d_fir4_nhwc(x, f, flip, gain, N, H, W, C, padx0, padx1, pady0, pady1, y, hi, lo, epi, stream):
        from oracle import restated as R
        xv = _f32(x, N, H, W, C).astype(np.float64)
        fv = _f32(f, 4, 4).astype(np.float32)
        out = R.upfirdn2d(np.ascontiguousarray(xv.transpose(0, 3, 1, 2)), fv, up=1, down=1, padding=[padx0, padx1, pady0, pady1], flip_filter=bool(flip), gain=gain)
        out = np.asarray(out, dtype=np.float64).transpose(0, 2, 3, 1)
        Ho, Wo = H + pady0 + pady1 - 3, W + padx0 + padx1 - 3
        assert out.shape == (N, Ho, Wo, C)
        if y:
            e = getattr(epi, '_obj', None)        # ctypes.byref(struct) keeps the structure in _obj; None = no epilogue
            if e is not None:
                out = _epilogue(out, e, N, Ho, Wo, C)
            _f32(y, N, Ho, Wo, C)[:] = out
        else:
            _store_bf16_pair(hi, lo, out, C)
        return 0

    @staticmethod
    def gp3d_demod_act_bwd_split(dy, y, d, noise, noise_scale, nps, b, dc, dc_hi, dc_lo, C_pad, g_d, g_b, g_ns, N, HW, C, act, alpha, gain, stream):
        dyv, yv = _f32(dy, N, HW, C).astype(np.float64), _f32(y, N, HW, C).astype(np.float64)
        slope = np.where(yv > 0, 1.0, alpha) if act == 3 else np.ones_like(yv)
        dt = dyv * gain * slope
        dv = _f32(d, N, 1, C).astype(np.float64) if d else np.ones([N, 1, C])
        dcv = dt * dv
        if g_b:
            _f32(g_b, C)[:] += dt.sum(axis=(0, 1))
        nz = None
        if noise:
            nz = (_f32(noise, N, HW, 1) if nps else _f32(noise, 1, HW, 1)).astype(np.float64)
            if g_ns:
                _f32(g_ns, 1)[:] += (dt * nz).sum()
        if g_d:      # the pre-demodulation conv value rebuilt from the saved output: y = act(c*d + noise*ns + b) * gain
            t = yv / gain / slope
            if nz is not None:
                t = t - nz * float(_f32(noise_scale, 1)[0])
            if b:
                t = t - _f32(b, 1, 1, C)
            _f32(g_d, N, C)[:] += (dt * (t / dv)).sum(axis=1)
        if dc:
            _f32(dc, N, HW, C)[:] = dcv
        else:
            _store_bf16_pair(dc_hi, dc_lo, dcv, C_pad)
        return 0

    @staticmethod
    def gp3d_act_bwd_split(dy, y, dc, dc_hi, dc_lo, C_pad, g_b, N, HW, C, act, alpha, gain, clamp, stream):
        alpha, gain, clamp = float(np.float32(alpha)), float(np.float32(gain)), float(np.float32(clamp))   # `float` parameters: a clipped y equals float32(clamp)
        dyv, yv = _f32(dy, N, HW, C).astype(np.float64), _f32(y, N, HW, C).astype(np.float64)
        slope = np.where(yv > 0, 1.0, alpha) if act == 3 else np.ones_like(yv)
        dt = dyv * gain * slope
        if clamp > 0:
            dt = np.where(np.abs(yv) < clamp, dt, 0.0)
        if g_b:
            _f32(g_b, C)[:] += dt.sum(axis=(0, 1))
        if dc:
            _f32(dc, N, HW, C)[:] = dt
        else:
            _store_bf16_pair(dc_hi, dc_lo, dt, C_pad)
        return 0

    @staticmethod
    def gp3d_modulate_bwd(dxs, x, s, dx, g_s, N, HW, C, stream):
        d, xv, sv = _f32(dxs, N, HW, C).astype(np.float64), _f32(x, N, HW, C).astype(np.float64), _f32(s, N, 1, C).astype(np.float64)
        _f32(dx, N, HW, C)[:] = d * sv
        _f32(g_s, N, C)[:] += (d * xv).sum(axis=1)
        return 0


# ---------------------------------------------------------------------------------------------------------------------------------
# Plugin-level stand-ins (the pybind11 surface of the reference: bias_act.cpp:94-97, upfirdn2d.cpp:102-105) for the Python wrappers
# 3dgp_b200/torch_utils/ops/{upfirdn2d, bias_act}.py -- their autograd Functions (adjoint padding, gradient-of-gradient chain) run unchanged on top.
class Upfirdn2dPlugin:
    @staticmethod
    def upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
        from oracle import restated as R
        y = R.upfirdn2d(x.detach().to(torch.float64).numpy(), f.detach().to(torch.float32).numpy(), up=[upx, upy], down=[downx, downy],
                        padding=[padx0, padx1, pady0, pady1], flip_filter=bool(flip), gain=gain)
        y = torch.from_numpy(np.ascontiguousarray(y)).to(x.dtype)
        return y.contiguous(memory_format=torch.channels_last) if (x.stride(1) == 1 and x.shape[1] > 1) else y


class BiasActPlugin:
    """bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp) for all nine activations and gradient orders 0 / 1 / 2, following the formula
    table of the reference kernel (bias_act.cu:23-147; restated in csrc/bias_act.cu::act_eval / bias_act_one), in float64."""

    @staticmethod
    def bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp):
        A, G = int(act), int(grad)
        alpha, gain, clamp = float(np.float32(alpha)), float(np.float32(gain)), float(np.float32(clamp))
        shape = [-1 if i == dim else 1 for i in range(x.dim())]
        f64 = lambda t: t.to(torch.float64) if t.numel() else None
        v, xr, yr, dyv = x.to(torch.float64), f64(xref), f64(yref), f64(dy)
        bb = b.to(torch.float64).reshape(shape) if b.numel() else 0.0
        if G == 0:
            v = v + bb
        elif xr is not None:
            xr = xr + bb
        yy = (yr / gain if gain != 0 else torch.zeros_like(yr)) if yr is not None else None
        ER, HER = 80.0, 40.0
        sS, sA = 1.0507009873554804934193349852946, 1.6732632423543772848170429916717
        W = torch.where
        if A == 1:
            y = v if G <= 1 else torch.zeros_like(v)
        elif A == 2:
            y = W(v > 0, v, torch.zeros_like(v)) if G == 0 else (W(yy > 0, v, torch.zeros_like(v)) if G == 1 else torch.zeros_like(v))
        elif A == 3:
            y = W(v > 0, v, v * alpha) if G == 0 else (W(yy > 0, v, v * alpha) if G == 1 else torch.zeros_like(v))
        elif A == 4:
            y = torch.tanh(v) if G == 0 else (v * (1 - yy * yy) if G == 1 else v * (1 - yy * yy) * (-2 * yy))
        elif A == 5:
            y = torch.sigmoid(v) if G == 0 else (v * yy * (1 - yy) if G == 1 else v * yy * (1 - yy) * (1 - 2 * yy))
        elif A == 6:
            y = W(v >= 0, v, torch.expm1(v.clamp(max=0))) if G == 0 else (W(yy >= 0, v, v * (yy + 1)) if G == 1 else W(yy >= 0, torch.zeros_like(v), v * (yy + 1)))
        elif A == 7:
            y = (W(v >= 0, sS * v, sS * sA * torch.expm1(v.clamp(max=0))) if G == 0 else
                 (W(yy >= 0, v * sS, v * (yy + sS * sA)) if G == 1 else W(yy >= 0, torch.zeros_like(v), v * (yy + sS * sA))))
        elif A == 8:
            if G == 0:
                y = W(v > ER, v, torch.log1p(torch.exp(v.clamp(max=ER))))
            else:
                c = torch.exp(-yy)
                y = v * (1 - c) if G == 1 else v * c * (1 - c)
        elif A == 9:
            if G == 0:
                y = v * torch.sigmoid(v)
            else:
                c = torch.exp(xr.clamp(max=ER)); d = c + 1
                y = W(xr > HER, v, v * c * (xr + d) / (d * d)) if G == 1 else W(xr > HER, torch.zeros_like(v), v * c * (xr * (2 - d) + 2 * d) / (d * d * d))
                yr = xr * torch.sigmoid(xr) * gain
        else:
            raise AssertionError(f'unknown activation index {A}')
        y = y * gain * (dyv if dyv is not None else 1.0)
        if clamp >= 0:
            y = y.clamp(-clamp, clamp) if G == 0 else W((yr > -clamp) & (yr < clamp), y, torch.zeros_like(y))
        out = torch.empty_like(x)
        out.copy_(y.to(x.dtype))
        return out


# ---------------------------------------------------------------------------------------------------------------------------------
# Fused ray-march entry points (gp3d_raymarch_forward / _forward_cam / _backward, gp3d_generate_rays) served by the oracle's renderer: the host wrapper
# 3dgp_b200/torch_utils/ops/raymarch.py (plane layout + strides, option codes, output shapes, gradient routing, camera chain) runs unchanged on top.
def _planes_view(ptr, B, C, P, psB, psP, psC, psY, psX):
    """[B, 3, C, P, P] float32 view of the plane tensor behind `ptr` with the element strides the ABI receives."""
    extent = (B - 1) * psB + 2 * psP + (C - 1) * psC + (P - 1) * psY + (P - 1) * psX + 1
    base = np.ctypeslib.as_array(ctypes.cast(int(ptr), ctypes.POINTER(ctypes.c_float)), shape=(extent,))
    return np.lib.stride_tricks.as_strided(base, shape=(B, 3, C, P, P), strides=tuple(4 * s for s in (psB, psP, psC, psY, psX)))


def _render_args(o, w1, b1, w2, b2, uc, uf, sc, sf):
    B, R, N, C, H = o.B, o.R, o.N, o.C, o.H
    t = lambda p, *s: torch.from_numpy(_f32(p, *s).copy())
    # variates that are not injected come from the kernel's Philox stream keyed by (seed, offset): here a numpy stream with the same key, so that the
    # backward launch regenerates what the forward drew (the VALUES differ from the kernel's; parity tests always inject)
    rs = np.random.RandomState([int(o.seed) & 0x7fffffff, int(o.offset) & 0x7fffffff])
    draw_u = lambda: torch.from_numpy(rs.uniform(0, 1, size=(B, R, N)).astype(np.float32))
    draw_n = lambda: torch.from_numpy(rs.standard_normal((B, R, N)).astype(np.float32))
    u_c = t(uc, B, R, N) if uc else draw_u()
    u_f = t(uf, B, R, N) if uf else draw_u()
    noisy = o.noise_std > 0
    s_c = t(sc, B, R, N) if sc else (draw_n() if noisy else None)
    s_f = t(sf, B, R, N) if sf else (draw_n() if noisy else None)
    return dict(w1=t(w1, H, C), b1=t(b1, H), w2=t(w2, 4, H), b2=t(b2, 4), u_coarse=u_c, u_fine=u_f, sn_coarse=s_c, sn_fine=s_f)


def _render(planes, a, ray_o, ray_d, o):
    from oracle import restated as R
    return R.render(planes, a['w1'], a['b1'], a['w2'], a['b2'], ray_o, ray_d, a['u_coarse'], a['u_fine'], o.ray_start, o.ray_end, o.box_half, o.N,
                    sn_coarse=a['sn_coarse'], sn_fine=a['sn_fine'], noise_std=o.noise_std, use_inf_depth=bool(o.use_inf_depth), last_back=bool(o.last_back),
                    white_back_end_idx=o.white_back_end_idx, clamp_mode={0: 'softplus', 1: 'relu'}[o.clamp_mode])


def _store_outputs(res, rgb, depth, wsum, tfin, B, R):
    _f32(rgb, B, R, 3)[:] = res[0].detach().numpy(); _f32(depth, B, R)[:] = res[1].detach().numpy()
    _f32(wsum, B, R)[:] = res[2].detach().numpy(); _f32(tfin, B, R)[:] = res[3].detach().numpy()


def _raymarch_forward(planes, planes_dtype, psB, psP, psC, psY, psX, ray_o, ray_d, w1, b1, w2, b2, uc, uf, sc, sf, rgb, depth, wsum, tfin, opts, stream):
    o = opts._obj
    assert planes_dtype == 0, 'float32 planes only in the emulation'
    pl = torch.from_numpy(_planes_view(planes, o.B, o.C, o.P, psB, psP, psC, psY, psX).copy())
    a = _render_args(o, w1, b1, w2, b2, uc, uf, sc, sf)
    ro, rd = torch.from_numpy(_f32(ray_o, o.B, o.R, 3).copy()), torch.from_numpy(_f32(ray_d, o.B, o.R, 3).copy())
    _store_outputs(_render(pl, a, ro, rd, o), rgb, depth, wsum, tfin, o.B, o.R)
    return 0


def _generate_rays(c2w, fov, ps, po, B, h, w, ray_o, ray_d, stream):
    from oracle import restated as R
    t = lambda p, *s: torch.from_numpy(_f32(p, *s).copy())
    ro, rd = R.sample_rays(t(c2w, B, 4, 4), t(fov, B), (w, h), t(ps, B, 2) if ps else None, t(po, B, 2) if po else None)
    _f32(ray_o, B, h * w, 3)[:] = ro.numpy(); _f32(ray_d, B, h * w, 3)[:] = rd.numpy()
    return 0


def _raymarch_forward_cam(planes, planes_dtype, psB, psP, psC, psY, psX, cam, w1, b1, w2, b2, uc, uf, sc, sf, rgb, depth, wsum, tfin, opts, stream):
    o, c = opts._obj, cam._obj
    R_ = c.img_h * c.img_w
    assert o.R == R_
    ro = torch.empty([o.B, R_, 3]); rd = torch.empty([o.B, R_, 3])
    _generate_rays(c.c2w, c.fov, c.patch_scales, c.patch_offsets, o.B, c.img_h, c.img_w, ro.data_ptr(), rd.data_ptr(), None)
    return _raymarch_forward(planes, planes_dtype, psB, psP, psC, psY, psX, ro.data_ptr(), rd.data_ptr(), w1, b1, w2, b2, uc, uf, sc, sf,
                             rgb, depth, wsum, tfin, opts, stream)


def _raymarch_backward(planes, planes_dtype, psB, psP, psC, psY, psX, ray_o, ray_d, w1, b1, w2, b2, uc, uf, sc, sf, g_rgb, g_depth,
                       g_planes, g_w1, g_b1, g_w2, g_b2, g_ray_o, g_ray_d, opts, stream):
    """Gradients are ACCUMULATED into the (zero-filled) MLP / plane buffers and WRITTEN to the ray buffers, as the kernel does."""
    o = opts._obj
    B, R_, C, P, H = o.B, o.R, o.C, o.P, o.H
    with torch.enable_grad():               # called from inside an autograd Function's backward, where grad mode is off
        pl = torch.from_numpy(_planes_view(planes, B, C, P, psB, psP, psC, psY, psX).copy()).requires_grad_(True)
        a = _render_args(o, w1, b1, w2, b2, uc, uf, sc, sf)
        for k in ('w1', 'b1', 'w2', 'b2'):
            a[k].requires_grad_(True)
        ro = torch.from_numpy(_f32(ray_o, B, R_, 3).copy()).requires_grad_(True); rd = torch.from_numpy(_f32(ray_d, B, R_, 3).copy()).requires_grad_(True)
        res = _render(pl, a, ro, rd, o)
        gr, gd = torch.from_numpy(_f32(g_rgb, B, R_, 3).copy()), torch.from_numpy(_f32(g_depth, B, R_).copy())
        grads = torch.autograd.grad([res[0], res[1]], [pl, a['w1'], a['b1'], a['w2'], a['b2'], ro, rd], [gr, gd])
    _planes_view(g_planes, B, C, P, psB, psP, psC, psY, psX)[...] += grads[0].numpy()
    _f32(g_w1, H, C)[:] += grads[1].numpy(); _f32(g_b1, H)[:] += grads[2].numpy(); _f32(g_w2, 4, H)[:] += grads[3].numpy(); _f32(g_b2, 4)[:] += grads[4].numpy()
    if g_ray_o:
        _f32(g_ray_o, B, R_, 3)[:] = grads[5].numpy(); _f32(g_ray_d, B, R_, 3)[:] = grads[6].numpy()
    return 0


FakeLib.gp3d_raymarch_forward = staticmethod(_raymarch_forward)
FakeLib.gp3d_raymarch_forward_cam = staticmethod(_raymarch_forward_cam)
FakeLib.gp3d_raymarch_backward = staticmethod(_raymarch_backward)
FakeLib.gp3d_generate_rays = staticmethod(_generate_rays)


# ---------------------------------------------------------------------------------------------------------------------------------
def install(monkeypatch, package='3dgp_b200'):
    """Swaps the whole emulation in: the ctypes library, the tensor-core launchers, the two plugins, and the CUDA-only guards of the public wrappers
    (each wrapper is entered right below its guard, at the autograd Function it dispatches to).  Eligibility predicates keep their shape rules and lose
    only the `is_cuda` clause, so the routing (fused node / tensor-core primitive / ATen) is the one a GPU run takes."""
    import contextlib
    import importlib
    _lib = importlib.import_module(package + '._lib')
    tc = importlib.import_module(package + '.torch_utils.ops.tc')
    modconv = importlib.import_module(package + '.torch_utils.ops.modconv')
    upf = importlib.import_module(package + '.torch_utils.ops.upfirdn2d')
    bact = importlib.import_module(package + '.torch_utils.ops.bias_act')
    gradfix = importlib.import_module(package + '.torch_utils.ops.conv2d_gradfix')
    from oracle import restated as R
    fake = FakeLib()
    monkeypatch.setattr(_lib, 'lib', lambda: fake)
    monkeypatch.setattr(_lib, 'stream_ptr', lambda: None)
    monkeypatch.setattr(_lib, 'require_cuda', lambda t, name='tensor': None)
    monkeypatch.setattr(torch.cuda, 'device', lambda _d: contextlib.nullcontext())
    monkeypatch.setattr(tc, 'split_bf16', split)
    monkeypatch.setattr(tc, 'conv_launch', conv_launch_epi)
    monkeypatch.setattr(tc, 'conv_transpose_s2_launch', conv_transpose_s2_launch)
    monkeypatch.setattr(tc, 'wgrad_launch', wgrad_launch)
    monkeypatch.setattr(upf, '_plugin', Upfirdn2dPlugin); monkeypatch.setattr(upf, '_init', lambda: True)
    monkeypatch.setattr(bact, '_plugin', BiasActPlugin); monkeypatch.setattr(bact, '_init', lambda: True)
    monkeypatch.setattr(upf, 'upfirdn2d', lambda x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda':
                        upf._upfirdn2d_cuda(up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain).apply(x, f))
    monkeypatch.setattr(bact, 'bias_act', lambda x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda':
                        bact._bias_act_cuda(dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp).apply(x, b))
    t2 = gradfix._tuple2
    monkeypatch.setattr(gradfix, 'conv2d', lambda input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1:
                        gradfix._conv(False, weight.shape, t2(stride), t2(padding), (0, 0), t2(dilation), groups, gradfix._terms_for(input.dtype)).apply(input, weight, bias))
    monkeypatch.setattr(gradfix, 'conv_transpose2d', lambda input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1:
                        gradfix._conv(True, weight.shape, t2(stride), t2(padding), t2(output_padding), t2(dilation), groups, gradfix._terms_for(input.dtype)).apply(input, weight, bias))

    class _AsCuda:                       # a tensor stand-in for the predicates: same shape / dtype, claims to live on the GPU
        def __init__(self, t):
            self.is_cuda, self.dtype, self.shape = True, t.dtype, t.shape
    el, cel = modconv.eligible, modconv.conv_act_eligible
    monkeypatch.setattr(modconv, 'eligible', lambda x, *a, **k: el(_AsCuda(x), *a, **k))
    monkeypatch.setattr(modconv, 'conv_act_eligible', lambda x, *a, **k: cel(_AsCuda(x), *a, **k))
    monkeypatch.setattr(R, 'DIFFERENTIABLE', True)
    return tc


class FilteredLreluPlugin:
    """filtered_lrelu_plugin (filtered_lrelu.cpp:16-298) with NO specialised kernel: `filtered_lrelu` answers return code -1, so the Python wrapper takes
    the reference's generic route upfirdn2d -> filtered_lrelu_act_ -> upfirdn2d (filtered_lrelu.py:223-229); `filtered_lrelu_act_` applies the
    sign-coded leaky ReLU in place and writes / reads the packed 2-bit sign tensor (0 positive, 1 negative, 2 clamped; four codes per byte along x,
    element (x, y) at sign coordinate (x + sx, y + sy): filtered_lrelu.cu:1136-1145)."""

    @staticmethod
    def filtered_lrelu(x, fu, fd, b, si, up, down, px0, px1, py0, py1, sx, sy, gain, slope, clamp, flip_filters, writeSigns):
        return torch.empty([0], dtype=x.dtype), torch.empty([0], dtype=torch.uint8), -1

    @staticmethod
    def filtered_lrelu_act_(x, si, sx, sy, gain, slope, clamp, writeSigns):
        N, C, H, W = x.shape
        gain, slope = float(np.float32(gain)), float(np.float32(slope))
        clamp = float(np.float32(clamp)) if (clamp is not None and clamp >= 0 and clamp != float('inf')) else -1.0
        v = x.detach().to(torch.float64).numpy()
        if writeSigns:
            code = np.zeros(v.shape, dtype=np.uint8)
            neg = v < 0
            v = np.where(neg, v * slope, v) * gain
            code[neg] = 1
            if clamp >= 0:
                big = np.abs(v) > clamp
                v = np.where(big, np.copysign(clamp, v), v)
                code[big] = 2
            sw = (W + 15) & ~15
            full = np.zeros([N, C, H, sw], dtype=np.uint8); full[..., :W] = code
            packed = (full[..., 0::4] | (full[..., 1::4] << 2) | (full[..., 2::4] << 4) | (full[..., 3::4] << 6)).astype(np.uint8)
            so = torch.from_numpy(packed)
        elif si.numel() > 0:
            s = si.numpy()
            sh, sw = s.shape[2], s.shape[3] * 4
            dec = np.stack([(s >> (2 * k)) & 3 for k in range(4)], axis=-1).reshape(N, C, sh, sw)
            ys, xs = np.arange(H) + sy, np.arange(W) + sx
            code = np.zeros([N, C, H, W], dtype=np.uint8)                           # outside the sign tensor: code 0 (positive)
            vy, vx = (ys >= 0) & (ys < sh), (xs >= 0) & (xs < sw)
            code[np.ix_(np.arange(N), np.arange(C), np.nonzero(vy)[0], np.nonzero(vx)[0])] = dec[np.ix_(np.arange(N), np.arange(C), ys[vy], xs[vx])]
            v = v * np.where(code == 0, gain, np.where(code == 1, gain * slope, 0.0))
            so = si
        else:
            v = np.where(v < 0, v * slope, v) * gain
            if clamp >= 0:
                v = np.clip(v, -clamp, clamp)
            so = si
        x.copy_(torch.from_numpy(v).to(x.dtype))
        return so


def _to_uint8(x, y, N, Cx, Cy, H, W, sN, sC, sH, sW, scale, shift, stream):
    """gp3d_to_uint8: y[n, c, h, w] = (uint8) clamp(x[n, c, h, w] * scale + shift, 0, 255) for the first Cy channels, x addressed by element strides."""
    extent = (N - 1) * sN + (Cx - 1) * sC + (H - 1) * sH + (W - 1) * sW + 1
    base = np.ctypeslib.as_array(ctypes.cast(int(x), ctypes.POINTER(ctypes.c_float)), shape=(extent,))
    xv = np.lib.stride_tricks.as_strided(base, shape=(N, Cx, H, W), strides=(4 * sN, 4 * sC, 4 * sH, 4 * sW))[:, :Cy]
    v = np.clip(xv * np.float32(scale) + np.float32(shift), 0, 255).astype(np.uint8)          # float32 arithmetic and a truncating cast, like torch's
    np.ctypeslib.as_array(ctypes.cast(int(y), ctypes.POINTER(ctypes.c_uint8)), shape=(N * Cy * H * W,))[:] = v.reshape(-1)
    return 0


FakeLib.gp3d_to_uint8 = staticmethod(_to_uint8)


# ---------------------------------------------------------------------------------------------------------------------------------
# LIBRARY-level stand-ins for the entry points behind the three reference plugins (gp3d_bias_act, gp3d_upfirdn2d(_out_size), gp3d_filtered_lrelu(_act)),
# served from host pointers as include/gp3d_b200.h words them.  With these, the product's plugin OBJECTS (3dgp_b200/torch_utils/custom_ops.py: argument
# checks, stride / size marshalling, output allocation) run on CPU tensors -- underneath the product's wrappers or, in tests/test_cpu_reference_wrappers.py,
# underneath the UNMODIFIED reference wrappers (integration level A of INTEGRATION.md).
_CT = {0: (ctypes.c_float, np.float32), 1: (ctypes.c_uint16, np.float16)}


def _flat(ptr, code, n):
    ct, npt = _CT[code]
    return np.ctypeslib.as_array(ctypes.cast(int(ptr), ctypes.POINTER(ct)), shape=(int(n),)).view(npt)


def _strided(ptr, code, shape, strides):
    extent = sum((d - 1) * s for d, s in zip(shape, strides)) + 1
    base = _flat(ptr, code, extent)
    return np.lib.stride_tricks.as_strided(base, shape=tuple(shape), strides=tuple(int(s) * base.itemsize for s in strides))


def _lib_bias_act(x, b, xref, yref, dy, y, dtype, numel, sizeB, stepB, grad, act, alpha, gain, clamp, stream):
    t = lambda p: torch.from_numpy(_flat(p, dtype, numel).copy()) if p else torch.empty([0], dtype=torch.float32 if dtype == 0 else torch.float16)
    xt = t(x)
    if b:       # the bias of element i is b[(i / stepB) % sizeB]
        bb = torch.from_numpy(_flat(b, dtype, sizeB).copy())[(torch.arange(numel) // stepB) % sizeB]
    else:
        bb = torch.empty([0], dtype=xt.dtype)
    out = BiasActPlugin.bias_act(xt, bb, t(xref), t(yref), t(dy), grad, 0, act, alpha, gain, clamp)
    _flat(y, dtype, numel)[:] = out.numpy()
    return 0


def _lib_upfirdn2d_out_size(in_size, up, down, pad0, pad1, fsize):
    from oracle import restated as R
    return R.upfirdn2d_out_size(in_size, up, down, pad0, pad1, fsize)


def _lib_upfirdn2d(x, f, y, dtype, N, C, inH, inW, xsN, xsC, xsH, xsW, fh, fw, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain,
                   outH, outW, ysN, ysC, ysH, ysW, stream):
    from oracle import restated as R
    xv = _strided(x, dtype, (N, C, inH, inW), (xsN, xsC, xsH, xsW)).astype(np.float64)
    fv = _flat(f, 0, fh * fw).reshape(fh, fw).copy()
    out = R.upfirdn2d(xv, fv, up=[upx, upy], down=[downx, downy], padding=[padx0, padx1, pady0, pady1], flip_filter=bool(flip), gain=gain)
    assert out.shape == (N, C, outH, outW), 'the caller-computed output extent disagrees with the operator'
    yv = _strided(y, dtype, (N, C, outH, outW), (ysN, ysC, ysH, ysW))
    yv[...] = out.astype(yv.dtype)
    return 0


def _lib_filtered_lrelu(*args):
    return -2       # GP3D_E_UNSUPPORTED: no fused kernel in the emulation, the caller takes the generic route


def _lib_filtered_lrelu_act(x, si, dtype, N, C, H, W, sH, sW4, sx, sy, gain, slope, clamp, write_signs, stream):
    xv = _strided(x, dtype, (N, C, H, W), (C * H * W, H * W, W, 1))
    xt = torch.from_numpy(xv.copy())
    if write_signs:
        assert sx == 0 and sy == 0 and sH == H and sW4 == ((W + 15) & ~15) >> 2
        so = FilteredLreluPlugin.filtered_lrelu_act_(xt, torch.empty([0], dtype=torch.uint8), 0, 0, gain, slope, clamp, True)
        np.ctypeslib.as_array(ctypes.cast(int(si), ctypes.POINTER(ctypes.c_uint8)), shape=(N, C, sH, sW4))[...] = so.numpy()
    else:
        st = torch.empty([0], dtype=torch.uint8)
        if si:
            st = torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(int(si), ctypes.POINTER(ctypes.c_uint8)), shape=(N, C, sH, sW4)).copy())
        FilteredLreluPlugin.filtered_lrelu_act_(xt, st, sx, sy, gain, slope, clamp, False)
    xv[...] = xt.numpy()
    return 0


def install_plugin_library(monkeypatch):
    """install() plus: the product's REAL plugin objects (custom_ops.get_plugin) on the library-level stand-ins above.  Returns the product's custom_ops."""
    import importlib
    install(monkeypatch)
    co = importlib.import_module('3dgp_b200.torch_utils.custom_ops')
    for name, fn in (('gp3d_bias_act', _lib_bias_act), ('gp3d_upfirdn2d_out_size', _lib_upfirdn2d_out_size), ('gp3d_upfirdn2d', _lib_upfirdn2d),
                     ('gp3d_filtered_lrelu', _lib_filtered_lrelu), ('gp3d_filtered_lrelu_act', _lib_filtered_lrelu_act)):
        monkeypatch.setattr(FakeLib, name, staticmethod(fn), raising=False)
    monkeypatch.setattr(FakeLib, 'gp3d_last_error', staticmethod(lambda: b'emulated'), raising=False)
    monkeypatch.setattr(co, '_x_on_cuda', lambda x: None)
    monkeypatch.setattr(co, '_cached_plugins', dict())
    return co
