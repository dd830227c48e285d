"""Filtered leaky ReLU: bias -> upsample FIR -> gain * lrelu -> clamp -> downsample FIR (StyleGAN3's anti-aliased
non-linearity).  Python surface of the reference op (src/torch_utils/ops/filtered_lrelu.py:56-272).

The reference ships this op but nothing on the 3DGP path calls it; BASELINE configs[3] names it.  Separable filters on
contiguous float32 / float16 tensors run as ONE fused kernel (csrc/filtered_lrelu.cu: the up-sampled intermediate never leaves shared
memory); everything else takes the reference's own *generic* route -- upfirdn2d, the in-place sign-coded activation kernel
`filtered_lrelu_act_`, upfirdn2d -- with every stage on lib3dgp_b200's sm_100a kernels.  Only the 2-bit/element sign tensor is retained
for backward (0 positive, 1 negative, 2 clamped; filtered_lrelu.cu:1136-1145), and the backward is the same op with up/down swapped
and the signs read back (filtered_lrelu.py:252-263).
"""
import numpy as np
import torch

from .. import custom_ops
from . import upfirdn2d
from .upfirdn2d import _parse_padding, _get_filter_size

_plugin = None


def _init():
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin(module_name='filtered_lrelu_plugin', sources=['misc.cu'], headers=['common.cuh'],
                                        source_dir=None, extra_cuda_cflags=['--use_fast_math'])
    return True


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None,
                   flip_filter=False, impl='cuda'):
    assert isinstance(x, torch.Tensor)
    if impl != 'cuda' or x.device.type != 'cuda':
        raise RuntimeError("3dgp_b200.filtered_lrelu has only the sm_100a implementation (impl='cuda', CUDA tensors)")
    _init()
    return _filtered_lrelu_cuda(up=up, down=down, padding=padding, gain=gain, slope=slope, clamp=clamp,
                                flip_filter=flip_filter).apply(x, fu, fd, b, None, 0, 0)


_cache = dict()


def _filtered_lrelu_cuda(up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None, flip_filter=False):
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1
    px0, px1, py0, py1 = _parse_padding(padding)
    gain = float(gain); slope = float(slope)
    assert gain > 0 and slope >= 0 and (clamp is None or clamp >= 0)
    clamp = float(clamp if clamp is not None else 'inf')
    key = (up, down, px0, px1, py0, py1, gain, slope, clamp, flip_filter)
    if key in _cache:
        return _cache[key]

    class FilteredLRelu(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, fu, fd, b, si, sx, sy):
            assert isinstance(x, torch.Tensor) and x.ndim == 4
            one = lambda: torch.ones([1, 1], dtype=torch.float32, device=x.device)
            fu = one() if fu is None else fu
            fd = one() if fd is None else fd
            assert 1 <= fu.ndim <= 2 and 1 <= fd.ndim <= 2
            if up == 1 and fu.ndim == 1 and fu.shape[0] == 1:
                fu = fu.square()[None]
            if down == 1 and fd.ndim == 1 and fd.shape[0] == 1:
                fd = fd.square()[None]
            if si is None:
                si = torch.empty([0], dtype=torch.uint8, device=x.device)
            write_signs = (si.numel() == 0) and (x.requires_grad or (b is not None and b.requires_grad))
            y = so = None
            return_code = -1
            if x.dtype in (torch.float16, torch.float32):
                y, so, return_code = _plugin.filtered_lrelu(x, fu, fd, b, si, up, down, px0, px1, py0, py1, sx, sy, gain, slope, clamp, flip_filter, write_signs)
            if return_code < 0:     # generic route (filtered_lrelu.py:223-229)
                y = x if b is None else x + b.reshape(1, -1, 1, 1)
                y = upfirdn2d.upfirdn2d(x=y, f=fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
                y = y.contiguous()
                if y.data_ptr() == x.data_ptr():
                    y = y.clone()
                so = _plugin.filtered_lrelu_act_(y, si, sx, sy, gain, slope, clamp, write_signs)   # in place on y
                y = upfirdn2d.upfirdn2d(x=y, f=fd, down=down, flip_filter=flip_filter)
            ctx.save_for_backward(fu, fd, (si if si.numel() else so))
            ctx.x_shape, ctx.y_shape, ctx.s_ofs = x.shape, y.shape, (sx, sy)
            return y

        @staticmethod
        def backward(ctx, dy):
            fu, fd, si = ctx.saved_tensors
            _, _, xh, xw = ctx.x_shape
            _, _, yh, yw = ctx.y_shape
            sx, sy = ctx.s_ofs
            dx = db = None
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[3]:
                pp = [(fu.shape[-1] - 1) + (fd.shape[-1] - 1) - px0, xw * up - yw * down + px0 - (up - 1),
                      (fu.shape[0] - 1) + (fd.shape[0] - 1) - py0, xh * up - yh * down + py0 - (up - 1)]
                gg = gain * (up ** 2) / (down ** 2)
                sx2 = sx - (fu.shape[-1] - 1) + px0
                sy2 = sy - (fu.shape[0] - 1) + py0
                dx = _filtered_lrelu_cuda(up=down, down=up, padding=pp, gain=gg, slope=slope, clamp=None,
                                          flip_filter=(not flip_filter)).apply(dy, fd, fu, None, si, sx2, sy2)
            if ctx.needs_input_grad[3]:
                db = dx.sum([0, 2, 3])
            return dx, None, None, db, None, None, None

    _cache[key] = FilteredLRelu
    return FilteredLRelu
