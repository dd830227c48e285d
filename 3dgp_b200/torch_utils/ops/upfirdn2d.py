"""Pad / upsample / FIR-filter / downsample of 2D images -- Python surface of the reference op
(src/torch_utils/ops/upfirdn2d.py:70-387) over the sm_100a kernels in csrc/upfirdn2d.cu.

The op is linear and its adjoint is the same op with up<->down swapped, the filter flipped and the padding
`p` below (upfirdn2d.py:250-269), so gradients of any order run on the same kernel.
"""
import collections

import numpy as np
import torch

from .. import custom_ops

_plugin = None


def _init():
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin(module_name='upfirdn2d_plugin', sources=['upfirdn2d.cu'], headers=['common.cuh'],
                                        source_dir=None, extra_cuda_cflags=['--use_fast_math'])
    return True


def _pair(v, what):
    v = [v, v] if isinstance(v, int) else v
    assert isinstance(v, (list, tuple)) and all(isinstance(e, int) for e in v), what
    return list(v)


def _parse_scaling(scaling):
    """int or [x, y] -> (x, y), both >= 1."""
    sx, sy = _pair(scaling, 'scaling')
    assert min(sx, sy) >= 1
    return sx, sy


def _parse_padding(padding):
    """int, [x, y] or [x0, x1, y0, y1] -> (x0, x1, y0, y1); negative entries crop."""
    p = _pair(padding, 'padding')
    if len(p) == 2:
        p = [p[0], p[0], p[1], p[1]]
    x0, x1, y0, y1 = p
    return x0, x1, y0, y1


def _get_filter_size(f):
    """(taps along x, taps along y); None is the identity filter."""
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    assert min(f.shape) >= 1
    return int(f.shape[-1]), int(f.shape[0])


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """FIR filter preparation: outer product for short 1-D filters (< 8 taps), DC normalisation, optional flip,
    gain ** (ndim/2).  Same results as the reference helper (upfirdn2d.py:70-114)."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in [0, 1, 2] and f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = f.ger(f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """See the reference docstring (upfirdn2d.py:118-157).  CUDA only."""
    assert isinstance(x, torch.Tensor)
    if impl != 'cuda' or x.device.type != 'cuda':
        raise RuntimeError("3dgp_b200.upfirdn2d has only the sm_100a implementation (impl='cuda', CUDA tensors)")
    _init()
    return _upfirdn2d_cuda(up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain).apply(x, f)


class FirSpec(collections.namedtuple('FirSpec', 'up down pad flip gain')):
    """Static half of one upfirdn2d call: up / down = (x, y) factors, pad = (x0, x1, y0, y1), filter orientation, gain.
    The operator is linear, and its adjoint is again an upfirdn2d -- `adjoint()` gives that call's spec -- so gradients of every order
    are more launches of the same kernel (the reference builds one autograd class per parameter set for this, upfirdn2d.py:214-273;
    here the spec travels as a plain argument of ONE autograd node)."""
    __slots__ = ()

    def out_extent(self, ih, iw, fh, fw):
        (ux, uy), (dx, dy), (x0, x1, y0, y1) = self.up, self.down, self.pad
        return (ih * uy + y0 + y1 - fh + dy) // dy, (iw * ux + x0 + x1 - fw + dx) // dx

    def adjoint(self, in_hw, out_hw, f_hw):
        """Spec of the transposed operator for an input of extent in_hw that produced out_hw with an f_hw filter."""
        (ux, uy), (dx, dy), (x0, _, y0, _) = self.up, self.down, self.pad
        (ih, iw), (oh, ow), (fh, fw) = in_hw, out_hw, f_hw
        lead_x, lead_y = fw - x0 - 1, fh - y0 - 1
        # the trailing edge restores exactly the input extent: whatever the forward decimation dropped comes back as zeros
        tail_x = iw * ux - ow * dx + x0 - ux + 1
        tail_y = ih * uy - oh * dy + y0 - uy + 1
        return FirSpec(self.down, self.up, (lead_x, tail_x, lead_y, tail_y), not self.flip, self.gain)

    def launch(self, x, f):
        """One or two kernel launches: a 2-D filter in one pass, a separable one as a row pass and a column pass (upfirdn2d.py:243-245)."""
        (ux, uy), (dx, dy), (x0, x1, y0, y1) = self.up, self.down, self.pad
        if f.ndim == 2:
            return _plugin.upfirdn2d(x, f, ux, uy, dx, dy, x0, x1, y0, y1, self.flip, self.gain)
        rows = _plugin.upfirdn2d(x, f.unsqueeze(0), ux, 1, dx, 1, x0, x1, 0, 0, self.flip, 1.0)
        return _plugin.upfirdn2d(rows, f.unsqueeze(1), 1, uy, 1, dy, 0, 0, y0, y1, self.flip, self.gain)

    def apply(self, x, f):
        return _FirNode.apply(x, f, self)


class _FirNode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, f, spec):
        assert isinstance(x, torch.Tensor) and x.ndim == 4
        if f is None:
            f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
        if f.ndim == 1 and f.shape[0] == 1:
            f = f.square().unsqueeze(0)   # a separable single tap is the full 1x1 filter f*f
        assert f.ndim in [1, 2]
        y = spec.launch(x, f)
        ctx.save_for_backward(f)
        ctx.spec, ctx.in_hw = spec, tuple(x.shape[2:])
        return y

    @staticmethod
    def backward(ctx, dy):
        f, = ctx.saved_tensors
        assert not ctx.needs_input_grad[1], 'the filter is a constant'
        if not ctx.needs_input_grad[0]:
            return None, None, None
        fw, fh = _get_filter_size(f)
        return ctx.spec.adjoint(ctx.in_hw, tuple(dy.shape[2:]), (fh, fw)).apply(dy, f), None, None


def _upfirdn2d_cuda(up=1, down=1, padding=0, flip_filter=False, gain=1):
    """Object with `.apply(x, f)` for one parameter set (the reference's private entry point of the same name, upfirdn2d.py:199-273)."""
    return FirSpec(_parse_scaling(up), _parse_scaling(down), _parse_padding(padding), bool(flip_filter), gain)


def _margins(taps, factor, upsampling):
    """(leading, trailing) padding that centres a `taps`-wide filter on a lattice `factor` times finer (upsampling) or coarser."""
    if upsampling:
        return (taps + factor - 1) // 2, (taps - factor) // 2
    return (taps - factor + 1) // 2, (taps - factor) // 2


def _centred(f, padding, fx=1, fy=1, upsampling=False):
    """User padding plus the centring margins of filter `f` for factors (fx, fy): the [x0, x1, y0, y1] list upfirdn2d takes."""
    fw, fh = _get_filter_size(f)
    (mx0, mx1), (my0, my1) = _margins(fw, fx, upsampling), _margins(fh, fy, upsampling)
    x0, x1, y0, y1 = _parse_padding(padding)
    return [x0 + mx0, x1 + mx1, y0 + my0, y1 + my1]


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """'same'-size FIR filtering (reference upfirdn2d.py:277-309): the factor-1 down-sampling margins, taps // 2 and (taps - 1) // 2."""
    return upfirdn2d(x, f, padding=_centred(f, padding), flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Zero-insert upsampling + FIR, output = input * up; the gain makes up for the inserted zeros (reference upfirdn2d.py:313-348)."""
    ux, uy = _parse_scaling(up)
    return upfirdn2d(x, f, up=up, padding=_centred(f, padding, ux, uy, upsampling=True), flip_filter=flip_filter, gain=gain * ux * uy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """FIR + decimation, output = input / down (reference upfirdn2d.py:352-387)."""
    dx, dy = _parse_scaling(down)
    return upfirdn2d(x, f, down=down, padding=_centred(f, padding, dx, dy), flip_filter=flip_filter, gain=gain, impl=impl)
