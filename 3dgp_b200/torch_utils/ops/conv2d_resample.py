"""2D convolution with optional FIR up/down-sampling, as a two-step "plan, then run" operator.

`conv2d_resample(x, w, f, up, down, padding, groups, flip_weight, flip_filter)` has the reference's public contract
(src/torch_utils/ops/conv2d_resample.py:46-141): same output extents, same FIR footprint per layer, same order of
the filter and the contraction, because those decide which pixels every layer of G and D sees.

The route is a pure function of a handful of integers (kernel / filter extents, the factors, the padding), so it is
computed once per layer shape (`plan`, memoised) as a short list of stages and then executed: every `Fir` stage is one
launch of csrc/upfirdn2d.cu / fir_tma.cu, every `Contract` stage one tap-list launch of the tcgen05 kernels through
conv2d_gradfix (a stride-`up` transposed contraction becomes the four polyphase tap lists there).
"""
import collections
import functools

import torch

from . import conv2d_gradfix
from . import upfirdn2d
from .upfirdn2d import _parse_padding, _get_filter_size

# One FIR launch: `filtered` False means pad / crop only (no filter taps);  pad = (x0, x1, y0, y1).
Fir = collections.namedtuple('Fir', 'filtered up down pad gain')
# One dense contraction: `mirrored` asks for the 180-degree rotated kernel;  pad = (y, x), symmetric.
Contract = collections.namedtuple('Contract', 'stride pad transposed mirrored')


def _filter_margins(taps, factor, upsampling):
    """Centring margins of the resampling filter; none when the factor is 1 (no filter is applied on that side)."""
    return upfirdn2d._margins(taps, factor, upsampling) if factor > 1 else (0, 0)


@functools.lru_cache(maxsize=None)
def plan(kh, kw, fw, fh, up, down, padding, flip_weight):
    """Stage list for one layer shape.  `padding` = (x0, x1, y0, y1) as the caller gave it, relative to the up-sampled lattice."""
    ux, dx = _filter_margins(fw, up, True), _filter_margins(fw, down, False)
    uy, dy = _filter_margins(fh, up, True), _filter_margins(fh, down, False)
    x0, x1 = padding[0] + ux[0] + dx[0], padding[1] + ux[1] + dx[1]
    y0, y1 = padding[2] + uy[0] + dy[0], padding[3] + uy[1] + dy[1]
    pointwise = (kh == 1 and kw == 1)
    turn = (kh > 1 or kw > 1)                       # a 1x1 kernel is its own mirror image
    dense = Contract(stride=1, pad=(0, 0), transposed=False, mirrored=(turn and not flip_weight))

    if up == 1 and down > 1:
        if pointwise:                               # decimate first: a quarter of the pixels reach the contraction
            return (Fir(True, 1, down, (x0, x1, y0, y1), 1), dense)
        return (Fir(True, 1, 1, (x0, x1, y0, y1), 1), dense._replace(stride=down))

    if up > 1 and down == 1 and pointwise:          # contract on the coarse lattice, then interpolate
        return (dense, Fir(True, up, 1, (x0, x1, y0, y1), up ** 2))

    if up > 1:
        # zero-insertion + kernel == stride-`up` transposed contraction; what is left of the padding moves into the FIR that follows,
        # and as much cropping as both edges share is done by the transposed contraction itself
        x0, x1 = x0 - (kw - 1), x1 - (kw - up)
        y0, y1 = y0 - (kh - 1), y1 - (kh - up)
        crop_x = max(min(-x0, -x1), 0)
        crop_y = max(min(-y0, -y1), 0)
        stages = [Contract(stride=up, pad=(crop_y, crop_x), transposed=True, mirrored=(turn and flip_weight)),
                  Fir(True, 1, 1, (x0 + crop_x, x1 + crop_x, y0 + crop_y, y1 + crop_y), up ** 2)]
        if down > 1:
            stages.append(Fir(True, 1, down, (0, 0, 0, 0), 1))
        return tuple(stages)

    # same resolution in and out
    if x0 == x1 and y0 == y1 and x0 >= 0 and y0 >= 0:
        return (dense._replace(pad=(y0, x0)),)
    return (Fir(False, 1, 1, (x0, x1, y0, y1), 1), dense)     # ragged / negative padding: explicit pad-or-crop pass


def _as_transposed_weight(w, groups):
    """[Cout, Cin/g, kh, kw] -> the [Cin, Cout/g, kh, kw] layout a transposed contraction reads."""
    if groups == 1:
        return w.transpose(0, 1)
    co, ci, kh, kw = w.shape
    return w.reshape(groups, co // groups, ci, kh, kw).transpose(1, 2).reshape(groups * ci, co // groups, kh, kw)


def _conv2d_wrapper(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """The reference's private helper, kept for callers that import it (conv2d_resample.py:20-42): correlation when `flip_weight`."""
    stage = Contract(stride, padding, transpose, (not flip_weight) and (w.shape[2] > 1 or w.shape[3] > 1))
    return _contract(x, w, stage, groups)


def _contract(x, w, stage, groups):
    if stage.mirrored:
        w = w.flip([2, 3])
    if stage.transposed:
        return conv2d_gradfix.conv_transpose2d(x, w, stride=stage.stride, padding=stage.pad, groups=groups)
    return conv2d_gradfix.conv2d(x, w, stride=stage.stride, padding=stage.pad, groups=groups)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert isinstance(w, torch.Tensor) and w.ndim == 4 and w.dtype == x.dtype
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in [1, 2] and f.dtype == torch.float32)
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1
    assert isinstance(groups, int) and groups >= 1
    fw, fh = _get_filter_size(f)
    stages = plan(int(w.shape[2]), int(w.shape[3]), fw, fh, up, down, tuple(_parse_padding(padding)), bool(flip_weight))
    for stage in stages:
        if isinstance(stage, Fir):
            x = upfirdn2d.upfirdn2d(x=x, f=(f if stage.filtered else None), up=stage.up, down=stage.down, padding=list(stage.pad),
                                    gain=stage.gain, flip_filter=flip_filter)
        else:
            x = _contract(x, _as_transposed_weight(w, groups) if stage.transposed else w, stage, groups)
    return x
