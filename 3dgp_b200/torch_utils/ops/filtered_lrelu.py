"""Filtered leaky ReLU: bias -> upsample FIR -> gain * lrelu -> clamp -> downsample FIR (StyleGAN3's anti-aliased
non-linearity).  Python surface of the reference op (src/torch_utils/ops/filtered_lrelu.py:56-272).

The reference ships this op but nothing on the 3DGP path calls it; BASELINE configs[3] names it.  Separable filters on
contiguous float32 / float16 tensors run as ONE fused kernel (csrc/filtered_lrelu.cu: the up-sampled intermediate never leaves shared
memory); everything else takes the reference's own *generic* route -- upfirdn2d, the in-place sign-coded activation kernel
`filtered_lrelu_act_`, upfirdn2d -- with every stage on lib3dgp_b200's sm_100a kernels.  Only the 2-bit/element sign tensor is retained
for backward (0 positive, 1 negative, 2 clamped; filtered_lrelu.cu:1136-1145), and the backward is the same op with up/down swapped
and the signs read back (filtered_lrelu.py:252-263).
"""
import collections

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


class LreluSpec(collections.namedtuple('LreluSpec', 'up down pad gain slope clamp flip')):
    """Static half of one filtered_lrelu call (pad = (x0, x1, y0, y1) on the up-sampled lattice, clamp = inf when off).  The backward pass is the
    same operator with the filters' roles exchanged, no clamp, and the stored sign codes read back at a shifted origin (`backward_spec`), so one
    autograd node parameterised by the spec serves every order (the reference builds a class per parameter set, filtered_lrelu.py:178-272)."""
    __slots__ = ()

    def backward_spec(self, x_hw, y_hw, fu, fd):
        """(spec, sign-origin shift) of the gradient call for an x_hw input that produced y_hw."""
        (xh, xw), (yh, yw) = x_hw, y_hw
        x0, _, y0, _ = self.pad
        reach_x = (fu.shape[-1] - 1) + (fd.shape[-1] - 1)          # taps either filter hangs over an edge
        reach_y = (fu.shape[0] - 1) + (fd.shape[0] - 1)
        pad = (reach_x - x0, xw * self.up - yw * self.down + x0 - (self.up - 1),
               reach_y - y0, xh * self.up - yh * self.down + y0 - (self.up - 1))
        gain = self.gain * (self.up ** 2) / (self.down ** 2)
        shift = (x0 - (fu.shape[-1] - 1), y0 - (fu.shape[0] - 1))
        return LreluSpec(self.down, self.up, pad, gain, self.slope, float('inf'), not self.flip), shift

    def run(self, x, fu, fd, b, signs, sx, sy, write_signs):
        """(y, signs-out).  The fused kernel when the plugin has one for this case; else bias -> up FIR -> sign-coded lrelu in place -> down FIR."""
        x0, x1, y0, y1 = self.pad
        if x.dtype in (torch.float16, torch.float32):
            y, so, rc = _plugin.filtered_lrelu(x, fu, fd, b, signs, self.up, self.down, x0, x1, y0, y1, sx, sy, self.gain, self.slope, self.clamp,
                                               self.flip, write_signs)
            if rc >= 0:
                return y, so
        t = x if b is None else x + b.reshape(1, -1, 1, 1)
        t = upfirdn2d.upfirdn2d(x=t, f=fu, up=self.up, padding=[x0, x1, y0, y1], gain=self.up ** 2, flip_filter=self.flip).contiguous()
        if t.data_ptr() == x.data_ptr():        # the activation kernel works in place: never on the caller's tensor
            t = t.clone()
        so = _plugin.filtered_lrelu_act_(t, signs, sx, sy, self.gain, self.slope, self.clamp, write_signs)
        return upfirdn2d.upfirdn2d(x=t, f=fd, down=self.down, flip_filter=self.flip), so

    def apply(self, x, fu, fd, b, si, sx, sy):
        return _FilteredLreluNode.apply(x, fu, fd, b, si, sx, sy, self)


def _unit_filter(f, factor, device):
    """None -> 1x1 identity; a separable single tap with no resampling -> the 1x1 filter f*f."""
    if f is None:
        return torch.ones([1, 1], dtype=torch.float32, device=device)
    assert 1 <= f.ndim <= 2
    if factor == 1 and f.ndim == 1 and f.shape[0] == 1:
        return f.square()[None]
    return f


class _FilteredLreluNode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fu, fd, b, si, sx, sy, spec):
        assert isinstance(x, torch.Tensor) and x.ndim == 4
        fu, fd = _unit_filter(fu, spec.up, x.device), _unit_filter(fd, spec.down, x.device)
        if si is None:
            si = torch.empty([0], dtype=torch.uint8, device=x.device)
        # sign codes are produced only by a first-order forward that will be differentiated; a gradient call reads the ones it was given
        write_signs = (si.numel() == 0) and (x.requires_grad or (b is not None and b.requires_grad))
        y, so = spec.run(x, fu, fd, b, si, sx, sy, write_signs)
        ctx.save_for_backward(fu, fd, (si if si.numel() else so))
        ctx.spec, ctx.x_hw, ctx.y_hw, ctx.s_ofs = spec, tuple(x.shape[2:]), tuple(y.shape[2:]), (sx, sy)
        return y

    @staticmethod
    def backward(ctx, dy):
        fu, fd, si = ctx.saved_tensors
        dx = db = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[3]:
            spec, (shx, shy) = ctx.spec.backward_spec(ctx.x_hw, ctx.y_hw, fu, fd)
            dx = spec.apply(dy, fd, fu, None, si, ctx.s_ofs[0] + shx, ctx.s_ofs[1] + shy)
        if ctx.needs_input_grad[3]:
            db = dx.sum([0, 2, 3])
        return dx, None, None, db, None, None, None, None


def _filtered_lrelu_cuda(up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None, flip_filter=False):
    """Object with `.apply(x, fu, fd, b, si, sx, sy)` (the reference's private entry point of the same name, filtered_lrelu.py:159-272)."""
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1
    gain, slope = float(gain), float(slope)
    assert gain > 0 and slope >= 0 and (clamp is None or clamp >= 0)
    return LreluSpec(up, down, tuple(_parse_padding(padding)), gain, slope, float('inf' if clamp is None else clamp), bool(flip_filter))
