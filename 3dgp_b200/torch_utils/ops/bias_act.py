"""Fused bias + activation (+gain, +clamp) -- same Python surface as the reference op
(src/torch_utils/ops/bias_act.py:52-86) on top of the sm_100a kernel in csrc/bias_act.cu.

First and second order gradients are supported through the same grad=1 / grad=2 kernel forms the reference uses
(bias_act.py:142-207).  There is no `impl='ref'` branch in the product: CPU tensors raise.
"""
import collections

import numpy as np
import torch

from ... import dnnlib
from .. import custom_ops

activation_funcs = {
    'linear':   dnnlib.EasyDict(def_alpha=0,   def_gain=1,          cuda_idx=1, ref='',  has_2nd_grad=False),
    'relu':     dnnlib.EasyDict(def_alpha=0,   def_gain=np.sqrt(2), cuda_idx=2, ref='y', has_2nd_grad=False),
    'lrelu':    dnnlib.EasyDict(def_alpha=0.2, def_gain=np.sqrt(2), cuda_idx=3, ref='y', has_2nd_grad=False),
    'tanh':     dnnlib.EasyDict(def_alpha=0,   def_gain=1,          cuda_idx=4, ref='y', has_2nd_grad=True),
    'sigmoid':  dnnlib.EasyDict(def_alpha=0,   def_gain=1,          cuda_idx=5, ref='y', has_2nd_grad=True),
    'elu':      dnnlib.EasyDict(def_alpha=0,   def_gain=1,          cuda_idx=6, ref='y', has_2nd_grad=True),
    'selu':     dnnlib.EasyDict(def_alpha=0,   def_gain=1,          cuda_idx=7, ref='y', has_2nd_grad=True),
    'softplus': dnnlib.EasyDict(def_alpha=0,   def_gain=1,          cuda_idx=8, ref='y', has_2nd_grad=True),
    'swish':    dnnlib.EasyDict(def_alpha=0,   def_gain=np.sqrt(2), cuda_idx=9, ref='x', has_2nd_grad=True),
}

_plugin = None


def _init():
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin(module_name='bias_act_plugin', sources=['bias_act.cu'], headers=['common.cuh'],
                                        source_dir=None, extra_cuda_cflags=['--use_fast_math'])
    return True


def _null(x):
    return torch.empty([0], dtype=x.dtype, device=x.device)


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    """y = clamp(act(x + b) * gain); see the reference docstring (bias_act.py:53-81)."""
    assert isinstance(x, torch.Tensor)
    if impl != 'cuda' or x.device.type != 'cuda':
        raise RuntimeError("3dgp_b200.bias_act has only the sm_100a implementation (impl='cuda', CUDA tensors)")
    _init()
    return _bias_act_cuda(dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp).apply(x, b)


class ActSpec(collections.namedtuple('ActSpec', 'dim act alpha gain clamp')):
    """Static half of one bias_act call, with defaults resolved (clamp = -1 when off).  The kernel has three forms selected by `order`:
    0: y = clamp(act(x + b) * gain);  1: dx from dy;  2: the second-order term d(dx)/dx.  Forms 1 and 2 read x or y -- whichever the
    activation's derivative is cheapest to express in (`ref` column of the table) -- so only those are kept for backward."""
    __slots__ = ()

    @property
    def entry(self):
        return activation_funcs[self.act]

    @property
    def is_identity(self):
        return self.act == 'linear' and self.gain == 1 and self.clamp < 0

    def kernel(self, t, b, x, y, dy, order):
        return _plugin.bias_act(t, b, x, y, dy, order, self.dim, self.entry.cuda_idx, self.alpha, self.gain, self.clamp)

    def reduce_to_bias(self, g):
        return g.sum([i for i in range(g.ndim) if i != self.dim])

    def apply(self, x, b):
        return _BiasActNode.apply(x, b, self)


def _layout_of(t):
    """Channel-minor 4-D tensors stay channel-minor through the op and its gradients; everything else is made dense row-major."""
    return torch.channels_last if (t.ndim == 4 and t.stride(1) == 1 and t.shape[1] > 1) else torch.contiguous_format


class _BiasActNode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, b, spec):
        ctx.spec, ctx.layout = spec, _layout_of(x)
        x = x.contiguous(memory_format=ctx.layout)
        nul = _null(x)
        bias = nul if b is None else b.contiguous()
        y = x if (spec.is_identity and b is None) else spec.kernel(x, bias, nul, nul, nul, 0)
        e = spec.entry
        needs_x = ('x' in e.ref) or e.has_2nd_grad
        ctx.save_for_backward(x if needs_x else nul, bias if needs_x else nul, y if 'y' in e.ref else nul)
        return y

    @staticmethod
    def backward(ctx, dy):
        spec = ctx.spec
        x, b, y = ctx.saved_tensors
        dx = db = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            dy = dy.contiguous(memory_format=ctx.layout)
            dx = dy if spec.is_identity else _BiasActGradNode.apply(dy, x, b, y, spec)
        if ctx.needs_input_grad[1]:
            db = spec.reduce_to_bias(dx)
        return dx, db, None


class _BiasActGradNode(torch.autograd.Function):
    """dx = dy * act'(.) * gain (zero where clamped): linear in dy, so its gradient w.r.t. dy is itself; w.r.t. x (and b) the order-2 kernel form."""

    @staticmethod
    def forward(ctx, dy, x, b, y, spec):
        ctx.spec, ctx.layout = spec, _layout_of(dy)
        dx = spec.kernel(dy, b, x, y, _null(dy), 1)
        ctx.save_for_backward(dy if spec.entry.has_2nd_grad else _null(dy), x, b, y)
        return dx

    @staticmethod
    def backward(ctx, d_dx):
        spec = ctx.spec
        d_dx = d_dx.contiguous(memory_format=ctx.layout)
        dy, x, b, y = ctx.saved_tensors
        d_dy = d_x = d_b = None
        if ctx.needs_input_grad[0]:
            d_dy = _BiasActGradNode.apply(d_dx, x, b, y, spec)
        if spec.entry.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
            d_x = spec.kernel(d_dx, b, x, y, dy, 2)
            if ctx.needs_input_grad[2]:
                d_b = spec.reduce_to_bias(d_x)
        return d_dy, d_x, d_b, None, None


def _bias_act_cuda(dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """Object with `.apply(x, b)` (the reference's private entry point of the same name, bias_act.py:118-207)."""
    assert clamp is None or clamp >= 0
    e = activation_funcs[act]
    return ActSpec(dim, act, float(e.def_alpha if alpha is None else alpha), float(e.def_gain if gain is None else gain), float(-1 if clamp is None else clamp))
