"""Plugin shim -- the drop-in boundary of the reference (src/torch_utils/custom_ops.py:59-155).

The reference's `get_plugin(module_name, sources, headers, source_dir, **build_kwargs)` JIT-compiles pybind11
modules whose functions take torch tensors.  Here the same call returns an object with the SAME function names and
argument lists (bias_act.cpp:94-97, upfirdn2d.cpp:102-105, filtered_lrelu.cpp:294-298), implemented on top of the
prebuilt C ABI (include/gp3d_b200.h, lib3dgp_b200.so, sm_100a).  `sources`/`headers`/`build_kwargs` are accepted for
signature compatibility and ignored: all kernels live in one shared object that `__graft_entry__.build()` compiles
with `-gencode arch=compute_100a,code=sm_100a` (the reference blanks TORCH_CUDA_ARCH_LIST, custom_ops.py:91, which
would yield plain sm_100 without the tcgen05/TMEM feature set).

Contract kept from the reference: callee allocates outputs, inputs are borrowed, empty tensors mean
"absent", errors surface as RuntimeError, kernels are enqueued on torch's current CUDA stream of x's device.
"""
import torch

from .. import _lib

verbosity = 'brief'  # kept for train.py:49-50 (`custom_ops.verbosity = 'none'`)

_cached_plugins = dict()


def _x_on_cuda(x):
    """The plugins' own device check (TORCH_CHECK(x.is_cuda(), ...) of bias_act.cpp:38, upfirdn2d.cpp:19, filtered_lrelu.cpp:22,218)."""
    if not x.is_cuda:
        raise RuntimeError('x must reside on CUDA device')


def _dev(x):
    _lib.require_cuda(x, 'x')
    return torch.cuda.device(x.device)


class _BiasActPlugin:
    """bias_act_plugin (reference bias_act.cpp:32-97)."""

    @staticmethod
    def bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp):
        L = _lib.lib()
        _x_on_cuda(x)
        for name, t in (('b', b), ('xref', xref), ('yref', yref), ('dy', dy)):
            if t.numel() > 0:
                if t.dtype != x.dtype or t.device != x.device:
                    raise RuntimeError(f'{name} must have the same dtype and device as x')
                if name != 'b' and tuple(t.shape) != tuple(x.shape):
                    raise RuntimeError(f'{name} must have the same shape as x')
        if x.numel() > 2147483647:
            raise RuntimeError('x is too large')
        dense = x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last))
        if not dense:
            raise RuntimeError('x must be contiguous')
        if b.numel() > 0:
            if b.dim() != 1:
                raise RuntimeError('b must have rank 1')
            if not (0 <= dim < x.dim()):
                raise RuntimeError('dim is out of bounds')
            if b.shape[0] != x.shape[dim]:
                raise RuntimeError('b has wrong number of elements')
            if not b.is_contiguous():
                raise RuntimeError('b must be contiguous')
        if grad < 0:
            raise RuntimeError('grad must be non-negative')
        for name, t in (('xref', xref), ('yref', yref), ('dy', dy)):
            if t.numel() > 0 and t.stride() != x.stride():
                t = t.contiguous(memory_format=torch.channels_last) if (x.dim() == 4 and not x.is_contiguous()) else t.contiguous()
                if name == 'xref': xref = t
                elif name == 'yref': yref = t
                else: dy = t
        y = torch.empty_like(x)
        if x.numel() == 0:
            return y
        has_b = b.numel() > 0
        with _dev(x):
            rc = L.gp3d_bias_act(
                x.data_ptr(), b.data_ptr() if has_b else None,
                xref.data_ptr() if xref.numel() else None, yref.data_ptr() if yref.numel() else None,
                dy.data_ptr() if dy.numel() else None, y.data_ptr(), _lib.dtype_code(x), x.numel(),
                b.numel() if has_b else 1, x.stride(dim) if has_b else 1,
                int(grad), int(act), float(alpha), float(gain), float(clamp), _lib.stream_ptr())
        _lib.check(rc, 'bias_act')
        return y


class _Upfirdn2dPlugin:
    """upfirdn2d_plugin (reference upfirdn2d.cpp:16-105)."""

    @staticmethod
    def upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
        L = _lib.lib()
        _x_on_cuda(x)
        if f.device != x.device:
            raise RuntimeError('f must reside on the same device as x')
        if f.dtype != torch.float32:
            raise RuntimeError('f must be float32')
        if x.numel() == 0:
            raise RuntimeError('x has zero size')
        if f.numel() == 0:
            raise RuntimeError('f has zero size')
        if x.dim() != 4:
            raise RuntimeError('x must be rank 4')
        if f.dim() != 2:
            raise RuntimeError('f must be rank 2')
        if upx < 1 or upy < 1:
            raise RuntimeError('upsampling factor must be at least 1')
        if downx < 1 or downy < 1:
            raise RuntimeError('downsampling factor must be at least 1')
        N, C, inH, inW = x.shape
        fh, fw = f.shape
        outW = L.gp3d_upfirdn2d_out_size(inW, upx, downx, padx0, padx1, fw)
        outH = L.gp3d_upfirdn2d_out_size(inH, upy, downy, pady0, pady1, fh)
        if outW < 1 or outH < 1:
            raise RuntimeError('output must be at least 1x1')
        channels_last = x.stride(1) == 1 and C > 1
        y = torch.empty([N, C, outH, outW], dtype=x.dtype, device=x.device,
                        memory_format=torch.channels_last if channels_last else torch.contiguous_format)
        fc = f.contiguous()
        with _dev(x):
            rc = L.gp3d_upfirdn2d(
                x.data_ptr(), fc.data_ptr(), y.data_ptr(), _lib.dtype_code(x), N, C, inH, inW,
                x.stride(0), x.stride(1), x.stride(2), x.stride(3),
                fh, fw, int(upx), int(upy), int(downx), int(downy), int(padx0), int(padx1), int(pady0), int(pady1),
                1 if flip else 0, float(gain), outH, outW,
                y.stride(0), y.stride(1), y.stride(2), y.stride(3), _lib.stream_ptr())
        _lib.check(rc, 'upfirdn2d')
        return y


class _FilteredLreluPlugin:
    """filtered_lrelu_plugin (reference filtered_lrelu.cpp:16-298).

    `filtered_lrelu` runs the fused sm_100a kernel (csrc/filtered_lrelu.cu) for separable filters and contiguous NCHW float32 / float16 tensors;
    any other configuration answers return_code -1 ("no specialised kernel", filtered_lrelu.cpp:50-55), which makes the caller take the generic
    path upfirdn2d -> filtered_lrelu_act_ -> upfirdn2d exactly as the reference does outside its own kernel table (filtered_lrelu.py:223-229).
    """

    @staticmethod
    def filtered_lrelu(x, fu, fd, b, si, up, down, px0, px1, py0, py1, sx, sy, gain, slope, clamp, flip_filters, writeSigns):
        L = _lib.lib()
        none = lambda: (torch.empty([0], device=x.device, dtype=x.dtype), torch.empty([0], device=x.device, dtype=torch.uint8), -1)
        _x_on_cuda(x)
        if fu.dtype != torch.float32 or fd.dtype != torch.float32:
            raise RuntimeError('fu and fd must be float32')
        if x.dim() != 4:
            raise RuntimeError('x must be rank 4')
        if x.numel() == 0:
            raise RuntimeError('x is empty')
        if fu.dim() not in (1, 2) or fd.dim() not in (1, 2):
            raise RuntimeError('fu and fd must be rank 1 or 2')
        if up < 1 or down < 1:
            raise RuntimeError('up and down must be at least 1')
        if b is not None and (b.dim() != 1 or b.shape[0] != x.shape[1]):
            raise RuntimeError('b must be a vector with the same number of channels as x')
        if b is not None and b.dtype != x.dtype:
            raise RuntimeError('x and b must have the same dtype')
        sep = lambda f: f.reshape(1) if tuple(f.shape) == (1, 1) else f           # a 1 x 1 filter is its own separable form
        fu, fd = sep(fu), sep(fd)
        if fu.dim() != 1 or fd.dim() != 1 or x.dtype not in (torch.float32, torch.float16) or not x.is_contiguous():
            return none()
        N, C, xh, xw = x.shape
        fut, fdt = fu.shape[0] - 1, fd.shape[0] - 1
        cw, ch = xw * up + (px0 + px1) - fut, xh * up + (py0 + py1) - fut           # logical size of the up-sampled buffer (filtered_lrelu.cpp:66-70)
        if not (cw > fdt and ch > fdt):
            raise RuntimeError('upsampled buffer must be at least the size of downsampling filter')
        yw, yh = (cw - fdt + (down - 1)) // down, (ch - fdt + (down - 1)) // down
        if yw < 1 or yh < 1:
            raise RuntimeError('output must be at least 1x1')
        read = si is not None and si.numel() > 0
        so = torch.empty([0], device=x.device, dtype=torch.uint8)
        s = si
        sw_active = 0
        if writeSigns:
            sw_active = yw * down - (down - 1) + fdt                                 # filtered_lrelu.cpp:86-93
            sh = yh * down - (down - 1) + fdt
            sw = (sw_active + 15) & ~15
            s = so = torch.zeros([N, C, sh, sw >> 2], dtype=torch.uint8, device=x.device)
        elif read:
            if si.dtype != torch.uint8 or si.dim() != 4 or not si.is_contiguous():
                raise RuntimeError('signs must be a contiguous rank-4 uint8 tensor')
            if si.shape[0] != N or si.shape[1] != C:
                raise RuntimeError('signs must have same batch & channels as x')
            sw_active = si.shape[3] << 2
        has_s = writeSigns or read
        y = torch.empty([N, C, yh, yw], dtype=x.dtype, device=x.device)
        fuc, fdc = fu.contiguous(), fd.contiguous()
        bc = b.contiguous() if b is not None else None
        with _dev(x):
            rc = L.gp3d_filtered_lrelu(
                x.data_ptr(), fuc.data_ptr(), fdc.data_ptr(), _lib.ptr(bc), s.data_ptr() if has_s else None, y.data_ptr(), _lib.dtype_code(x),
                N, C, xh, xw, yh, yw, fuc.shape[0], fdc.shape[0], int(up), int(down), int(px0), int(py0),
                s.shape[2] if has_s else 0, s.shape[3] if has_s else 0, int(sx), int(sy), (sw_active + 3) >> 2,
                float(gain), float(slope), float(clamp) if clamp is not None and clamp != float('inf') else -1.0, 1 if flip_filters else 0,
                1 if writeSigns else 0, 1 if (read and not writeSigns) else 0, _lib.stream_ptr())
        if rc == _lib.E_UNSUPPORTED:
            return none()
        _lib.check(rc, 'filtered_lrelu')
        return y, so, 0

    @staticmethod
    def filtered_lrelu_act_(x, si, sx, sy, gain, slope, clamp, writeSigns):
        L = _lib.lib()
        _x_on_cuda(x)
        if x.dim() != 4:
            raise RuntimeError('x must be rank 4')
        if not x.is_contiguous():
            raise RuntimeError('x must be contiguous')
        N, C, H, W = x.shape
        so = si
        read = si.numel() > 0
        if writeSigns:
            sw = (W + 15) & ~15   # width padded to a multiple of 16 elements (filtered_lrelu.cpp:89-93)
            so = torch.zeros([N, C, H, sw >> 2], dtype=torch.uint8, device=x.device)
        elif read:
            if si.dtype != torch.uint8 or si.dim() != 4 or not si.is_contiguous():
                raise RuntimeError('signs must be a contiguous rank-4 uint8 tensor')
        has_s = writeSigns or read
        with _dev(x):
            rc = L.gp3d_filtered_lrelu_act(
                x.data_ptr(), so.data_ptr() if has_s else None, _lib.dtype_code(x), N, C, H, W,
                so.shape[2] if has_s else 0, so.shape[3] if has_s else 0, int(sx), int(sy),
                float(gain), float(slope), float(clamp if clamp is not None else -1.0), 1 if writeSigns else 0,
                _lib.stream_ptr())
        _lib.check(rc, 'filtered_lrelu_act_')
        return so


_PLUGINS = {
    'bias_act_plugin': _BiasActPlugin,
    'upfirdn2d_plugin': _Upfirdn2dPlugin,
    'filtered_lrelu_plugin': _FilteredLreluPlugin,
}


def get_plugin(module_name, sources=None, headers=None, source_dir=None, **build_kwargs):
    """Same signature as the reference (custom_ops.py:59).  Returns the sm_100a-backed plugin object."""
    if module_name in _cached_plugins:
        return _cached_plugins[module_name]
    if module_name not in _PLUGINS:
        raise RuntimeError(f'unknown plugin "{module_name}" (available: {sorted(_PLUGINS)})')
    _lib.lib()  # loads (or builds) lib3dgp_b200.so; raises if impossible -- there is no fallback
    plugin = _PLUGINS[module_name]()
    _cached_plugins[module_name] = plugin
    return plugin
