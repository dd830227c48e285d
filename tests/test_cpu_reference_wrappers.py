"""Integration level A of INTEGRATION.md, executed: the UNMODIFIED reference wrappers (src/torch_utils/ops/{bias_act, upfirdn2d, filtered_lrelu}.py, imported in
place from /root/reference) on top of THIS repo's `custom_ops.get_plugin` -- the one function through which the reference binds its native code
(src/torch_utils/custom_ops.py:59; callers bias_act.py:38-48, upfirdn2d.py:23-33, filtered_lrelu.py:23-33).

The reference's `_init()` is left to call `get_plugin(module_name=..., sources=..., headers=..., source_dir=..., extra_cuda_cflags=...)` itself; what comes back
are the product's real plugin objects (argument checks, stride / size marshalling, output allocation of 3dgp_b200/torch_utils/custom_ops.py), here on the
library-level emulation of the C ABI (tests/abi_emulator.py::install_plugin_library; there is no GPU in this container).  Results are compared with the
reference's own `impl='ref'` implementations on the same inputs: forward and the gradients the reference's autograd classes derive through the plugin
(first and second order for bias_act, the adjoint call for upfirdn2d, the sign-coded backward for filtered_lrelu).

Skipped where /root/reference is absent (the GPU box); the kernels behind the same entry points are held to the same goldens by tests/test_gpu_ops.py."""
import numpy as np
import pytest
import torch

import abi_emulator as emu
from oracle import cases, ref_harness
from util import maxrel

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason='the unmodified reference is only present in the build container')


@pytest.fixture()
def ref(monkeypatch):
    """Reference op modules whose `custom_ops.get_plugin` is the product's."""
    co = emu.install_plugin_library(monkeypatch)
    ns = ref_harness.load()
    import src.torch_utils.custom_ops as ref_co
    calls = []

    def get_plugin(*a, **k):
        calls.append(k.get('module_name', a[0] if a else None))
        return co.get_plugin(*a, **k)
    monkeypatch.setattr(ref_co, 'get_plugin', get_plugin)
    for m in (ns.bias_act, ns.upfirdn2d, ns.filtered_lrelu):
        monkeypatch.setattr(m, '_plugin', None)
        assert m._init()                                  # the reference's own call into get_plugin, with its own argument list
        assert type(m._plugin).__module__.endswith('custom_ops') and type(m._plugin).__module__.startswith('3dgp_b200')
    assert calls == ['bias_act_plugin', 'upfirdn2d_plugin', 'filtered_lrelu_plugin']
    # CPU tensors stand in for CUDA ones: the reference's `x.device.type == 'cuda'` dispatch (upfirdn2d.py:161-163) is taken for it, and its stream query
    # (filtered_lrelu.py:215) answers "default stream"
    up = ns.upfirdn2d
    monkeypatch.setattr(up, 'upfirdn2d', lambda x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda':
                        ns.upfirdn2d._upfirdn2d_cuda(up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain).apply(x, f))
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda device=None: 0)
    monkeypatch.setattr(torch.cuda, 'default_stream', lambda device=None: 0)
    return ns


@pytest.mark.parametrize('name,kw', cases.bias_act_cases(), ids=[c[0] for c in cases.bias_act_cases()])
def test_reference_bias_act_wrapper_on_our_plugin(ref, name, kw):
    ba = ref.bias_act
    x, b = cases.bias_act_inputs(name, kw)
    args = dict(dim=kw['dim'], act=kw['act'], alpha=kw.get('alpha'), gain=kw.get('gain'), clamp=kw.get('clamp'))

    def run(fn):
        xt = torch.from_numpy(x).requires_grad_(True)
        bt = torch.from_numpy(b).requires_grad_(True) if b is not None else None
        y = fn(xt, bt)
        dy = torch.from_numpy(cases.cotangent(y.shape, 11))
        gr = torch.autograd.grad(y, [xt] + ([bt] if bt is not None else []), dy, create_graph=True)
        g2 = None
        if gr[0].requires_grad:
            g2 = torch.autograd.grad(gr[0], xt, torch.from_numpy(cases.cotangent(y.shape, 12)), allow_unused=True)[0]
        return [y] + list(gr) + [g2 if g2 is not None else torch.zeros_like(xt)]
    got = run(lambda xt, bt: ba._bias_act_cuda(**args).apply(xt, bt))                 # what bias_act(impl='cuda') dispatches to on a CUDA tensor (bias_act.py:84-85)
    want = run(lambda xt, bt: ba._bias_act_ref(x=xt, b=bt, **args))
    for g, w in zip(got, want):
        assert np.abs(g.detach().numpy() - w.detach().numpy()).max() < 1e-5 * max(1.0, w.detach().abs().max().item())


@pytest.mark.parametrize('name,kw', cases.upfirdn2d_cases(), ids=[c[0] for c in cases.upfirdn2d_cases()])
@pytest.mark.parametrize('layout', ['nchw', 'channels_last'])
def test_reference_upfirdn2d_wrapper_on_our_plugin(ref, name, kw, layout):
    up = ref.upfirdn2d
    x, f = cases.upfirdn2d_inputs(name, kw)
    ft = None if f is None else torch.from_numpy(f)
    args = dict(up=kw['up'], down=kw['down'], padding=kw['padding'], flip_filter=kw['flip_filter'], gain=kw['gain'])

    def run(fn):
        xt = torch.from_numpy(x)
        if layout == 'channels_last':
            xt = xt.contiguous(memory_format=torch.channels_last)
        xt.requires_grad_(True)
        y = fn(xt)
        gx, = torch.autograd.grad(y, xt, torch.from_numpy(cases.cotangent(y.shape, 5)))
        return y, gx
    y, gx = run(lambda xt: up._upfirdn2d_cuda(**args).apply(xt, ft))
    yr, gxr = run(lambda xt: up._upfirdn2d_ref(xt, ft, **args))
    assert y.shape == yr.shape and maxrel(y.detach().numpy(), yr.detach().numpy()) < 1e-5
    assert maxrel(gx.numpy(), gxr.numpy()) < 1e-5
    if layout == 'channels_last' and x.shape[1] > 1:
        assert y.stride(1) == 1                           # the plugin keeps the layout of its input (upfirdn2d.cpp:38-39)


@pytest.mark.parametrize('name,kw', cases.filtered_lrelu_cases()[:5], ids=[c[0] for c in cases.filtered_lrelu_cases()[:5]])
def test_reference_filtered_lrelu_wrapper_on_our_plugin(ref, name, kw):
    fl = ref.filtered_lrelu
    x, fu, fd, b = cases.filtered_lrelu_inputs(name, kw)
    tf = lambda a: None if a is None else torch.from_numpy(a)
    args = dict(up=kw['up'], down=kw['down'], padding=kw['padding'], gain=kw['gain'], slope=kw['slope'], clamp=kw['clamp'], flip_filter=False)

    def run(fn):
        xt = torch.from_numpy(x).requires_grad_(True); bt = torch.from_numpy(b).requires_grad_(True)
        y = fn(xt, bt)
        return (y,) + torch.autograd.grad(y, [xt, bt], torch.from_numpy(cases.cotangent(y.shape, 13)))
    got = run(lambda xt, bt: fl._filtered_lrelu_cuda(**args).apply(xt, tf(fu), tf(fd), bt, None, 0, 0))   # filtered_lrelu.py:117
    want = run(lambda xt, bt: fl._filtered_lrelu_ref(xt, fu=tf(fu), fd=tf(fd), b=bt, **args))
    for g, w in zip(got, want):
        assert g.shape == w.shape and maxrel(g.detach().numpy(), w.detach().numpy()) < 2e-5


def test_plugin_argument_errors_are_the_reference_messages(ref):
    """The checks of the C++ plugins (bias_act.cpp:38-52, upfirdn2d.cpp:19-31) as RuntimeErrors from the plugin objects."""
    p = ref.bias_act._plugin
    x = torch.zeros(2, 3, 4, 4); e = torch.empty(0)
    with pytest.raises(RuntimeError, match='b must have rank 1'):
        p.bias_act(x, torch.zeros(3, 1), e, e, e, 0, 1, 3, 0.2, 1.0, -1.0)
    with pytest.raises(RuntimeError, match='wrong number of elements'):
        p.bias_act(x, torch.zeros(4), e, e, e, 0, 1, 3, 0.2, 1.0, -1.0)
    with pytest.raises(RuntimeError, match='same shape as x'):
        p.bias_act(x, e, torch.zeros(2, 3, 4, 5), e, e, 1, 1, 3, 0.2, 1.0, -1.0)
    q = ref.upfirdn2d._plugin
    with pytest.raises(RuntimeError, match='rank 2'):
        q.upfirdn2d(x, torch.ones(4), 1, 1, 1, 1, 0, 0, 0, 0, False, 1.0)
    with pytest.raises(RuntimeError, match='at least 1x1'):
        q.upfirdn2d(torch.zeros(1, 1, 2, 2), torch.ones(5, 5), 1, 1, 1, 1, 0, 0, 0, 0, False, 1.0)
