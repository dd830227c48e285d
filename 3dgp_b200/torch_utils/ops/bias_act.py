"""Fused bias + activation (+gain, +clamp) -- same Python surface as the reference op
(src/torch_utils/ops/bias_act.py:52-86) on top of the sm_100a kernel in csrc/bias_act.cu.

First and second order gradients are supported through the same grad=1 / grad=2 kernel forms the reference uses
(bias_act.py:142-207).  There is no `impl='ref'` branch in the product: CPU tensors raise.
"""
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


_cache = dict()


def _bias_act_cuda(dim=1, act='linear', alpha=None, gain=None, clamp=None):
    assert clamp is None or clamp >= 0
    spec = activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    key = (dim, act, alpha, gain, clamp)
    if key in _cache:
        return _cache[key]
    idx = spec.cuda_idx
    keep_x = ('x' in spec.ref) or spec.has_2nd_grad
    keep_y = 'y' in spec.ref
    trivial = (act == 'linear' and gain == 1 and clamp < 0)

    class BiasAct(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, b):
            ctx.memory_format = torch.channels_last if (x.ndim == 4 and x.stride(1) == 1 and x.shape[1] > 1) else torch.contiguous_format
            x = x.contiguous(memory_format=ctx.memory_format)
            nul = _null(x)
            bb = b.contiguous() if b is not None else nul
            y = x
            if not trivial or b is not None:
                y = _plugin.bias_act(x, bb, nul, nul, nul, 0, dim, idx, alpha, gain, clamp)
            ctx.save_for_backward(x if keep_x else nul, bb if keep_x else nul, y if keep_y else nul)
            return y

        @staticmethod
        def backward(ctx, dy):
            dy = dy.contiguous(memory_format=ctx.memory_format)
            x, b, y = ctx.saved_tensors
            dx = db = None
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                dx = dy if trivial else BiasActGrad.apply(dy, x, b, y)
            if ctx.needs_input_grad[1]:
                db = dx.sum([i for i in range(dx.ndim) if i != dim])
            return dx, db

    class BiasActGrad(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x, b, y):
            ctx.memory_format = torch.channels_last if (dy.ndim == 4 and dy.stride(1) == 1 and dy.shape[1] > 1) else torch.contiguous_format
            dx = _plugin.bias_act(dy, b, x, y, _null(dy), 1, dim, idx, alpha, gain, clamp)
            ctx.save_for_backward(dy if spec.has_2nd_grad else _null(dy), x, b, y)
            return dx

        @staticmethod
        def backward(ctx, d_dx):
            d_dx = d_dx.contiguous(memory_format=ctx.memory_format)
            dy, x, b, y = ctx.saved_tensors
            d_dy = d_x = d_b = None
            if ctx.needs_input_grad[0]:
                d_dy = BiasActGrad.apply(d_dx, x, b, y)
            if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
                d_x = _plugin.bias_act(d_dx, b, x, y, dy, 2, dim, idx, alpha, gain, clamp)
            if spec.has_2nd_grad and ctx.needs_input_grad[2]:
                d_b = d_x.sum([i for i in range(d_x.ndim) if i != dim])
            return d_dy, d_x, d_b, None

    _cache[key] = BiasAct
    return BiasAct
