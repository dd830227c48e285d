"""Tensor-core (tcgen05 / TMEM) contraction ops of lib3dgp_b200: NHWC implicit-GEMM convolutions and their weight gradients."""
import weakref

import torch

from ... import _lib


def conv2d_nhwc_bf16(x, w, out=None, accumulate=False):
    """Stride-1 'same' convolution (correlation) as an implicit GEMM.  x: [N,H,W,Cin] bf16 (NHWC contiguous),
    w: [Cout,kh,kw,Cin] bf16 with kh == kw in {1, 3}; returns float32 [N,H,W,Cout]."""
    L = _lib.lib()
    _lib.require_cuda(x, 'x')
    if x.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        raise RuntimeError('conv2d_nhwc_bf16: operands must be bfloat16')
    x = x.contiguous(); w = w.contiguous()
    N, H, W, Cin = x.shape
    Cout, kh, kw, Cin2 = w.shape
    if Cin != Cin2 or kh != kw:
        raise RuntimeError('conv2d_nhwc_bf16: bad weight shape')
    if out is None:
        out = torch.empty([N, H, W, Cout], dtype=torch.float32, device=x.device)
        accumulate = False
    with torch.cuda.device(x.device):
        rc = L.gp3d_conv2d_nhwc_bf16(x.data_ptr(), w.data_ptr(), out.data_ptr(), N, H, W, Cin, Cout, kh, 1 if accumulate else 0, _lib.stream_ptr())
    _lib.check(rc, 'conv2d_nhwc_bf16')
    return out


def split_bf16(x_nhwc, styles=None, want_lo=True, pad_to=None, fp16=False):
    """x [N, ..., C] channel-minor contiguous (float32 / float16) -> (hi, lo) bfloat16 with x * styles == hi + lo up to 2^-16.
    styles: optional float32 [N, C] per-sample channel scale fused into the split.  pad_to: output channel count (zero tail).
    fp16=True: float16 operand(s) -- one for the discriminator's fp16-class blocks, a pair (22 bits) for the two-term x2w16 form."""
    L = _lib.lib()
    _lib.require_cuda(x_nhwc, 'x')
    assert x_nhwc.is_contiguous()
    N = x_nhwc.shape[0]; C = x_nhwc.shape[-1]
    Cp = C if pad_to is None else int(pad_to)
    HW = x_nhwc.numel() // (N * C)
    hi = torch.empty(list(x_nhwc.shape[:-1]) + [Cp], dtype=torch.float16 if fp16 else torch.bfloat16, device=x_nhwc.device)
    lo = torch.empty_like(hi) if want_lo else None
    with torch.cuda.device(x_nhwc.device):
        rc = L.gp3d_split_pad(x_nhwc.data_ptr(), _lib.dtype_code(x_nhwc), _lib.ptr(styles), hi.data_ptr(), _lib.ptr(lo), N, HW, C, Cp, 1 if fp16 else 0, _lib.stream_ptr())
    _lib.check(rc, 'split')
    return hi, lo


# Precision codes of the tensor-core convolutions ("terms"):
#   3  bf16x3 : x = xh + xl, w = wh + wl (bf16 pairs), xh*wh + xh*wl + xl*wh            3 MMAs / product, ~2^-16, fp32-grade
#   2  x2w16  : (xh + xl) * w16, fp16 activation PAIR x ONE fp16 weight operand          2 MMAs / product, weight rounding 2^-12
#               (forward activations only; a product with a gradient operand -- unbounded range -- runs as bf16x3)
#   1  bf16   : single bf16 product                                                      1 MMA
#   16 f16    : single fp16 x fp16 product for activation x weight -- the arithmetic class of the reference's fp16 discriminator blocks (fp16
#               operands, fp32 accumulation), with fp32 storage; a product with a gradient operand (range) is bf16 x bf16
# Both operands of one MMA share their element format: tcgen05.mma.kind::f16 with a_format != b_format is an illegal instruction on sm_100a
# (measured: tools/probe_formats.py, profiles/r2_operand_format_probe.txt).
def effective_terms(terms, grad=False):
    """Precision code actually executed: products with a gradient operand (`grad`) leave the fp16 forms (range)."""
    if terms == 2:
        return 3 if grad else 2
    if terms == 16:
        return 1 if grad else 16
    return terms


def operand_formats(terms, x_is_grad=False, w_is_grad=False):
    """(activation has low half, weight has low half, activation is fp16, weight is fp16) for a precision code."""
    return {3: (True, True, False, False), 2: (True, False, True, True), 16: (False, False, True, True),
            1: (False, False, False, False)}[effective_terms(terms, x_is_grad or w_is_grad)]


_weight_cache = {}
stats = dict(tc=0, aten=0, fused=0)   # primitive convs on the tcgen05 kernels / on ATen, fused layer nodes (ops/modconv.py); read by tests and bench.py
CONV_TIMING = None   # set to a list to collect (start_event, end_event, algorithmic_flops) of every bf16x3 conv launch -- conv_nhwc_bf16_kernel<*, 3>, all tap forms (bench.py roofline)


def timed(flops, launch):
    """Runs `launch()`; when CONV_TIMING is a list, brackets it with CUDA events on the current stream."""
    if CONV_TIMING is None:
        return launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = launch()
    e1.record()
    CONV_TIMING.append((e0, e1, flops))
    return r


def invalidate_weight_cache(param_ids=None):
    """Drop cached weight operands (for writers that bypass autograd's version counters, e.g. the fused optimiser kernel):
    all of them, or those of the parameters whose id() is in `param_ids`."""
    if param_ids is None:
        _weight_cache.clear()
    else:
        for key in [k_ for k_ in _weight_cache if k_[0] in param_ids]:
            del _weight_cache[key]


def tag_weight_source(w, param, gain):
    """Marks `w` (= (param * gain) possibly cast) with the Parameter it was derived from, so that the conv wrappers below can reuse the
    bf16 operands of that parameter across the forward / input-gradient calls of one optimiser step (Conv2dLayer, layers.py:229)."""
    w._gp3d_src = (param, float(gain))
    return w


def _operands_of(w, tag, make_nhwc, terms, w_is_grad=False):
    """Operands of conv weight `w` in layout make_nhwc(w) for the (effective) precision `terms`: cached per source Parameter when `w` carries a tag."""
    src = getattr(w, '_gp3d_src', None)
    if src is not None and src[0].shape == w.shape and isinstance(src[0], torch.nn.Parameter) and not w_is_grad:
        param, gain = src
        dt = w.dtype
        return weight_operands(param, (tag, gain, dt), lambda p_: make_nhwc((p_ * gain).to(dt).to(torch.float32)), terms)
    return _make_weight_operands(make_nhwc(w.detach().to(torch.float32)).contiguous(), terms, None)


def _make_weight_operands(wn, terms, pad_to):
    _, w_lo, _, w_fp16 = operand_formats(terms)
    return split_bf16(wn, want_lo=w_lo, pad_to=pad_to, fp16=w_fp16)


def weight_operands(weight, tag, make_nhwc, terms=3, pad_to=None):
    """Operand(s) of a weight tensor in the layout `make_nhwc(weight)` produces ([rows][taps][cols], cols contiguous) for precision `terms`:
    bf16 (hi, lo) pair (3), bf16 (hi, None) (1) or fp16 (w16, None) (2, 16); `terms` is an EFFECTIVE code (effective_terms).
    Parameters are re-laid-out and split ONCE per optimiser step: the cache is keyed by (id, tag) and invalidated by the tensor's
    autograd version counter (bumped by every in-place update).  Every entry holds a weak reference to its parameter whose callback removes
    the entry: an id() reused by a new tensor after the old one died can never produce a hit, and no operand copy outlives its weight."""
    key = (id(weight), tag, terms, pad_to)
    ver = (weight._version, weight.data_ptr(), tuple(weight.shape))
    hit = _weight_cache.get(key)
    if hit is not None and hit[0] == ver and hit[3]() is weight:
        return hit[1], hit[2]
    wn = make_nhwc(weight.detach().to(torch.float32)).contiguous()
    wh, wl = _make_weight_operands(wn, terms, pad_to)
    if isinstance(weight, torch.nn.Parameter):
        # the entry dies with its parameter (discarded networks, deep-copied snapshots): no operand copies outlive their weights on the device
        _weight_cache[key] = (ver, wh, wl, weakref.ref(weight, lambda _r, key=key: _weight_cache.pop(key, None)))
    return wh, wl


def channels_eligible(Cin, Cout):
    return Cin % 64 == 0 and (Cout % 128 == 0 or Cout in (64, 96))


def conv_eligible(N, Cin, H, W, Cout, k, stride, padding, dilation, groups):
    """Stride-1 'same' shapes the tcgen05 implicit-GEMM conv covers (csrc/conv_tc.cu)."""
    if groups != 1 or tuple(stride) != (1, 1) or tuple(dilation) != (1, 1) or k not in (1, 3, 5) or tuple(padding) != (k // 2, k // 2):
        return False
    return channels_eligible(Cin, Cout)


def _prep(x, w, tag, make_nhwc, terms, x_is_grad=False, w_is_grad=False):
    terms = effective_terms(terms, x_is_grad or w_is_grad)
    x_lo, _, x_fp16, w_fp16 = operand_formats(terms)
    xn = x.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)          # NHWC view, contiguous
    xh, xl = split_bf16(xn, want_lo=x_lo, fp16=x_fp16)
    wh, wl = _operands_of(w, tag, make_nhwc, terms, w_is_grad)
    return xh, xl, wh, wl


def conv_launch(xh, xl, wh, wl, y, N, H, W, Cin, Cout, slabs, taps, in_stride, HoP, WoP, Hout, Wout, osy, osx, oy0, ox0, epi=None, what='conv_nhwc'):
    """One tap-convolution launch through the descriptor entry point (gp3d_conv_nhwc).  Operand formats are read off the tensors' dtypes; the
    precision follows from which low-order halves are present (include/gp3d_b200.h: gp3d_conv_desc).  Multi-term launches (the tri-plane decoder's
    fp32-grade convolutions: bench.py's dominant kernel) are bracketed by CUDA events when CONV_TIMING is a list."""
    import ctypes
    L = _lib.lib()
    arr = (ctypes.c_int * (3 * len(taps)))(*[v for t in taps for v in t])
    d = _lib.ConvDesc(xh.data_ptr(), _lib.ptr(xl), wh.data_ptr(), _lib.ptr(wl), 1 if wh.dtype == torch.float16 else 0, 1 if xh.dtype == torch.float16 else 0,
                      y.data_ptr(), N, H, W, Cin, Cout, slabs, len(taps), ctypes.cast(arr, ctypes.c_void_p), in_stride, HoP, WoP, Hout, Wout, osy, osx, oy0, ox0, 0,
                      ctypes.pointer(epi) if epi is not None else None)
    launch = lambda: L.gp3d_conv_nhwc(ctypes.byref(d), _lib.stream_ptr())
    with torch.cuda.device(y.device):
        rc = timed(2.0 * N * HoP * WoP * Cin * Cout * len(taps), launch) if xl is not None else launch()
    _lib.check(rc, what)


def conv_transpose_s2_launch(xh, xl, wh, wl, y, N, H, W, Cin, Cout):
    """Stride-2 transposed 3x3 conv (padding 0) -> y [N, 2H+1, 2W+1, Cout]: the four polyphase tap convolutions as phases of ONE launch
    (gp3d_conv_transpose_s2_nhwc).  wh / wl: [Cout][9][Cin]."""
    L = _lib.lib()
    launch = lambda: L.gp3d_conv_transpose_s2_nhwc(xh.data_ptr(), _lib.ptr(xl), wh.data_ptr(), _lib.ptr(wl), 1 if wh.dtype == torch.float16 else 0,
                                                   1 if xh.dtype == torch.float16 else 0, y.data_ptr(), N, H, W, Cin, Cout, _lib.stream_ptr())
    flops = 2.0 * N * Cin * Cout * ((H + 1) * (W + 1) + 2 * (H + 1) * W + 2 * H * (W + 1) + 4 * H * W)       # taps x domain of the four phases
    with torch.cuda.device(y.device):
        rc = timed(flops, launch) if xl is not None else launch()
    _lib.check(rc, 'conv_transpose_s2_nhwc')


def same_taps(k):
    return [(ky - k // 2, kx - k // 2, ky * k + kx) for ky in range(k) for kx in range(k)]


def _taps_launch(xh, xl, wh, wl, y, N, H, W, Cin, Cout, slabs, taps, in_stride, HoP, WoP, Hout, Wout, osy, osx, oy0, ox0):
    conv_launch(xh, xl, wh, wl, y, N, H, W, Cin, Cout, slabs, taps, in_stride, HoP, WoP, Hout, Wout, osy, osx, oy0, ox0, what='conv_taps_nhwc')


def conv2d_forward(x, w, terms, adjoint=False, x_is_grad=False, w_is_grad=False):
    """Stride-1 'same' conv (adjoint=True: w is a conv weight used flipped + transposed, i.e. the input-gradient form).  x [N,Cin,H,W] (any strides,
    float32/float16), w [Cout,Cin,k,k] -> y [N,Cout,H,W] in x.dtype, channels-last strides.  terms: precision code (see operand_formats)."""
    N, Cin, H, W = x.shape
    if adjoint:     # y = conv(x, flip(w).transpose(0, 1)): operand layout [w.shape[1]][k][k][w.shape[0]]
        Cout, k = w.shape[1], w.shape[2]
        xh, xl, wh, wl = _prep(x, w, 'adj1', lambda w_: w_.flip([2, 3]).permute(1, 2, 3, 0), terms, x_is_grad, w_is_grad)
    else:
        Cout, _, k, _ = w.shape
        xh, xl, wh, wl = _prep(x, w, 'fwd', lambda w_: w_.permute(0, 2, 3, 1), terms, x_is_grad, w_is_grad)
    y = torch.empty([N, H, W, Cout], dtype=torch.float32, device=x.device)
    conv_launch(xh, xl, wh, wl, y, N, H, W, Cin, Cout, k * k, same_taps(k), 1, H, W, H, W, 1, 1, 0, 0, what='conv2d_nhwc')
    y = y.permute(0, 3, 1, 2)
    return y if x.dtype == torch.float32 else y.to(x.dtype)


def conv2d_strided_forward(x, w, stride, padding, terms, x_is_grad=False, w_is_grad=False):
    """Stride-2 conv (correlation): y[n,co,i,j] = sum x[n,ci,2i+ky-p,2j+kx-p] w[co,ci,ky,kx]; one strided-gather launch."""
    assert stride == 2
    N, Cin, H, W = x.shape
    Cout, _, k, _ = w.shape
    Ho = (H + 2 * padding - k) // 2 + 1
    Wo = (W + 2 * padding - k) // 2 + 1
    xh, xl, wh, wl = _prep(x, w, 'fwd', lambda w_: w_.permute(0, 2, 3, 1), terms, x_is_grad, w_is_grad)
    y = torch.empty([N, Ho, Wo, Cout], dtype=torch.float32, device=x.device)
    taps = [(ky - padding, kx - padding, ky * k + kx) for ky in range(k) for kx in range(k)]
    _taps_launch(xh, xl, wh, wl, y, N, H, W, Cin, Cout, k * k, taps, 2, Ho, Wo, Ho, Wo, 1, 1, 0, 0)
    y = y.permute(0, 3, 1, 2)
    return y if x.dtype == torch.float32 else y.to(x.dtype)


def conv_transpose2d_s2_forward(x, w, output_padding, terms, x_is_grad=False, w_is_grad=False):
    """Stride-2 transposed conv, padding 0 (conv_transpose2d weight layout [Cin, Cout, k, k], k == 3):
    y[n,co,2i+ky,2j+kx] += x[n,ci,i,j] w[ci,co,ky,kx], evaluated as four polyphase tap-convolutions on the low-resolution grid."""
    N, Cin, H, W = x.shape
    _, Cout, k, _ = w.shape
    assert k == 3
    Hout, Wout = 2 * H + 1 + output_padding[0], 2 * W + 1 + output_padding[1]
    xh, xl, wh, wl = _prep(x, w, 'tr2', lambda w_: w_.permute(1, 2, 3, 0), terms, x_is_grad, w_is_grad)           # [Cout,3,3,Cin]
    alloc = torch.zeros if (output_padding[0] or output_padding[1]) else torch.empty
    y = alloc([N, Hout, Wout, Cout], dtype=torch.float32, device=x.device)
    if not (output_padding[0] or output_padding[1]):
        conv_transpose_s2_launch(xh, xl, wh, wl, y, N, H, W, Cin, Cout)
    else:       # a padded output tensor is larger than 2H+1: the four phases as separate lattice launches
        for a in (0, 1):
            kys = [(0, 0), (-1, 2)] if a == 0 else [(0, 1)]         # (input offset, ky)
            for b in (0, 1):
                kxs = [(0, 0), (-1, 2)] if b == 0 else [(0, 1)]
                taps = [(dy, dx, ky * 3 + kx) for (dy, ky) in kys for (dx, kx) in kxs]
                _taps_launch(xh, xl, wh, wl, y, N, H, W, Cin, Cout, 9, taps, 1, H + 1 - a, W + 1 - b, Hout, Wout, 2, 2, a, b)
    y = y.permute(0, 3, 1, 2)
    return y if x.dtype == torch.float32 else y.to(x.dtype)


def wgrad_eligible(Cin, Cout):
    return Cin % 64 == 0 and Cout % 64 == 0


def wgrad_launch(dh, dl, xh, xl, dW, N, Hd, Wd, Cy, Hx, Wx, Cx, slabs, taps, sa, sb, HoP, WoP):
    """Weight-gradient pixel GEMM (gp3d_wgrad_taps_nhwc_fmt); operand formats are read off the tensors' dtypes (bf16 pairs, or single bf16 / fp16)."""
    import ctypes
    L = _lib.lib()
    arr = (ctypes.c_int * (5 * len(taps)))(*[v for t_ in taps for v in t_])
    with torch.cuda.device(dW.device):
        rc = L.gp3d_wgrad_taps_nhwc_fmt(dh.data_ptr(), _lib.ptr(dl), xh.data_ptr(), _lib.ptr(xl), 1 if dh.dtype == torch.float16 else 0,
                                        1 if xh.dtype == torch.float16 else 0, dW.data_ptr(), N, Hd, Wd, Cy, Hx, Wx, Cx, slabs,
                                        len(taps), ctypes.cast(arr, ctypes.c_void_p), sa, sb, HoP, WoP, _lib.stream_ptr())
    _lib.check(rc, 'wgrad_taps_nhwc')


def conv_wgrad(dy, x, k, mode, stride, padding, terms, x_is_grad=False):
    """Weight gradient on the tcgen05 pixel-GEMM (csrc/wgrad_tc.cu).
    mode 'conv'      : y = conv2d(x, w[Cout,Cin,k,k], stride, padding)            -> returns dW [Cout,Cin,k,k]
    mode 'transpose' : y = conv_transpose2d(x, w[Cin,Cout,k,k], stride 2, pad 0)  -> returns dW [Cin,Cout,k,k]
    dy: gradient of y, x: the op's input (any strides; float32 / float16).  A weight gradient always has a gradient operand: bf16x3, or one
    bf16 x bf16 product for the single-term codes (effective_terms)."""
    terms = effective_terms(terms, True)
    N, Cx, Hx, Wx = x.shape
    _, Cy, Hy, Wy = dy.shape
    dn = dy.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)
    xn = x.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)
    dh, dl = split_bf16(dn, want_lo=(terms == 3))                                   # gradients stay bf16 (range)
    xh, xl = split_bf16(xn, want_lo=(terms == 3))
    if mode == 'conv':
        # M operand = dy (Cout = Cy), N operand = x (Cin = Cx); pixel domain = output grid
        taps = [(0, 0, ky - padding, kx - padding, ky * k + kx) for ky in range(k) for kx in range(k)]
        sa, sb, HoP, WoP = 1, stride, Hy, Wy
    else:
        # y1[2i+ky][2j+kx] += x[i][j] w[ci][co][ky][kx]: M operand = dy1 read at stride 2 offset (ky,kx), N operand = x; domain = input grid
        assert stride == 2 and padding == 0
        taps = [(ky, kx, 0, 0, ky * k + kx) for ky in range(k) for kx in range(k)]
        sa, sb, HoP, WoP = 2, 1, Hx, Wx
    dW = torch.zeros([Cy, k * k, Cx], dtype=torch.float32, device=x.device)
    wgrad_launch(dh, dl, xh, xl, dW, N, Hy, Wy, Cy, Hx, Wx, Cx, k * k, taps, sa, sb, HoP, WoP)
    dW = dW.view(Cy, k, k, Cx)
    out = dW.permute(0, 3, 1, 2) if mode == 'conv' else dW.permute(3, 0, 1, 2)
    return out.contiguous().to(x.dtype)
