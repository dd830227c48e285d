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
    x = (xh + (xl if xl is not None else 0)).numpy().reshape(N, H, W, Cin)
    w = (wh + (wl if wl is not None else 0)).numpy().reshape(Cout, slabs, Cin)
    assert tuple(y.shape) == (N, Hout, Wout, Cout)
    acc = np.zeros([N, HoP, WoP, Cout], dtype=np.float64)
    for (dy, dx, slab) in taps:
        assert 0 <= slab < slabs
        acc += _shifted(x, dy, dx, in_stride, HoP, WoP) @ w[:, slab, :].T
    assert oy0 + (HoP - 1) * osy < Hout and ox0 + (WoP - 1) * osx < Wout, 'output lattice leaves the tensor'
    y[:, oy0:oy0 + (HoP - 1) * osy + 1:osy, ox0:ox0 + (WoP - 1) * osx + 1:osx, :] = torch.from_numpy(acc).to(y.dtype)


def conv_transpose_s2_launch(xh, xl, wh, wl, y, N, H, W, Cin, Cout):
    """y[n][2i+ky][2j+kx][co] += x[n][i][j][ci] * w[co][ky*3+kx][ci], y [N][2H+1][2W+1][Cout] fully written  (gp3d_conv_transpose_s2_nhwc)."""
    x = (xh + (xl if xl is not None else 0)).numpy().reshape(N, H, W, Cin)
    w = (wh + (wl if wl is not None else 0)).numpy().reshape(Cout, 9, Cin)
    out = np.zeros([N, 2 * H + 1, 2 * W + 1, Cout], dtype=np.float64)
    for ky in range(3):
        for kx in range(3):
            out[:, ky:ky + 2 * H:2, kx:kx + 2 * W:2, :] += x @ w[:, ky * 3 + kx, :].T
    y.copy_(torch.from_numpy(out).to(y.dtype))


def wgrad_launch(dh, dl, xh, xl, dW, N, Hd, Wd, Cy, Hx, Wx, Cx, slabs, taps, sa, sb, HoP, WoP):
    """dW[co][slab_t][ci] += sum_{n,iy,ix} dy[n][iy*sa+ay_t][ix*sa+ax_t][co] * x[n][iy*sb+by_t][ix*sb+bx_t][ci]  (gp3d_wgrad_taps_nhwc_fmt)."""
    d = (dh + (dl if dl is not None else 0)).numpy().reshape(N, Hd, Wd, Cy)
    x = (xh + (xl if xl is not None else 0)).numpy().reshape(N, Hx, Wx, Cx)
    assert tuple(dW.shape) == (Cy, slabs, Cx)
    acc = dW.to(torch.float64).numpy().copy()
    for (ay, ax, by, bx, slab) in taps:
        a = _shifted(d, ay, ax, sa, HoP, WoP).reshape(-1, Cy)
        b = _shifted(x, by, bx, sb, HoP, WoP).reshape(-1, Cx)
        acc[:, slab, :] += a.T @ b
    dW.copy_(torch.from_numpy(acc).to(dW.dtype))
