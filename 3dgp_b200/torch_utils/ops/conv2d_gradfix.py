"""conv2d / conv_transpose2d with arbitrarily high order gradients and the `no_weight_gradients()` switch the loss
uses for R1 / path-length regularisation (reference src/torch_utils/ops/conv2d_gradfix.py:24-172).

Engine selection: shapes covered by the tcgen05 implicit-GEMM kernels of lib3dgp_b200 (csrc/conv_tc.cu) are routed
there by `..ops.modconv`; everything else (tiny 4x4..16x16 layers, 5x5 depth-adaptor convs, strided / transposed
forms) goes through ATen's convolution, exactly the library call the reference makes (conv2d_gradfix.py:113-115).
"""
import contextlib

import torch

from . import tc

enabled = False                     # kept for training_loop.py:78 (`conv2d_gradfix.enabled = True`)
weight_gradients_disabled = False   # toggled by no_weight_gradients()
tc_enabled = True                   # route eligible convs (and their input-gradient convs) to the tcgen05 kernels
tc_stats = tc.stats                 # how many primitive convolutions went where (bench.py / tests); same dict as ops.tc.stats


_forced_terms = None


@contextlib.contextmanager
def tc_terms(n):
    """Precision of the tensor-core convs issued inside the block (ops.tc.operand_formats): 3 = error-compensated bf16x3 (fp32-grade, default for
    float32 tensors), 16 = single product with fp16 activations / weights and bf16 gradients, fp32 accumulation and storage (the class of arithmetic
    the reference's fp16 discriminator blocks use), 1 = single bf16 product.
    The choice is captured at forward time and reused by the corresponding backward convs."""
    global _forced_terms
    old = _forced_terms
    _forced_terms = n
    try:
        yield
    finally:
        _forced_terms = old


def _terms_for(dtype):
    if _forced_terms is not None:
        return _forced_terms
    return 3 if dtype == torch.float32 else 1


def _primitive_conv(input, weight, bias, stride, padding, dilation, groups, terms, xg=False, wg=False):
    k = weight.shape[2]
    if (tc_enabled and bias is None and weight.shape[2] == weight.shape[3] and input.dtype in (torch.float32, torch.float16)
            and tc.conv_eligible(input.shape[0], input.shape[1], input.shape[2], input.shape[3], weight.shape[0], k, stride, padding, dilation, groups)):
        tc_stats['tc'] += 1
        return tc.conv2d_forward(input, weight, terms, x_is_grad=xg, w_is_grad=wg)
    if (tc_enabled and bias is None and k == 3 and weight.shape[3] == 3 and tuple(stride) == (2, 2) and padding[0] == padding[1]
            and tuple(dilation) == (1, 1) and groups == 1 and input.dtype in (torch.float32, torch.float16)
            and tc.channels_eligible(input.shape[1], weight.shape[0])):
        tc_stats['tc'] += 1
        return tc.conv2d_strided_forward(input, weight, 2, padding[0], terms, x_is_grad=xg, w_is_grad=wg)
    tc_stats['aten'] += 1
    return torch.nn.functional.conv2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding, dilation=dilation, groups=groups)


def _primitive_conv_transpose(input, weight, bias, stride, padding, output_padding, dilation, groups, terms, xg=False, wg=False):
    # stride-1 transposed conv == correlation with the flipped, transposed kernel and padding k-1-p
    k = weight.shape[2]
    if (tc_enabled and bias is None and tuple(stride) == (1, 1) and tuple(output_padding) == (0, 0) and weight.shape[2] == weight.shape[3]
            and tuple(padding) == (k - 1 - k // 2, k - 1 - k // 2) and input.dtype in (torch.float32, torch.float16)
            and tc.conv_eligible(input.shape[0], input.shape[1], input.shape[2], input.shape[3], weight.shape[1], k, (1, 1), (k // 2, k // 2), dilation, groups)):
        tc_stats['tc'] += 1
        return tc.conv2d_forward(input, weight, terms, adjoint=True, x_is_grad=xg, w_is_grad=wg)
    if (tc_enabled and bias is None and k == 3 and weight.shape[3] == 3 and tuple(stride) == (2, 2) and tuple(padding) == (0, 0)
            and tuple(dilation) == (1, 1) and groups == 1 and input.dtype in (torch.float32, torch.float16)
            and tc.channels_eligible(input.shape[1], weight.shape[1])):
        tc_stats['tc'] += 1
        return tc.conv_transpose2d_s2_forward(input, weight, tuple(output_padding), terms, x_is_grad=xg, w_is_grad=wg)
    tc_stats['aten'] += 1
    return torch.nn.functional.conv_transpose2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding,
                                                output_padding=output_padding, groups=groups, dilation=dilation)


@contextlib.contextmanager
def no_weight_gradients(disable=True):
    global weight_gradients_disabled
    old = weight_gradients_disabled
    if disable:
        weight_gradients_disabled = True
    try:
        yield
    finally:        # an exception inside the block must not leave weight gradients off for the rest of the run
        weight_gradients_disabled = old


def _tuple2(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    if input.device.type != 'cuda':
        raise RuntimeError('3dgp_b200.conv2d_gradfix: CUDA tensors only (no CPU path)')
    return _conv(False, weight.shape, _tuple2(stride), _tuple2(padding), (0, 0), _tuple2(dilation), groups, _terms_for(input.dtype)).apply(input, weight, bias)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    if input.device.type != 'cuda':
        raise RuntimeError('3dgp_b200.conv2d_gradfix: CUDA tensors only (no CPU path)')
    return _conv(True, weight.shape, _tuple2(stride), _tuple2(padding), _tuple2(output_padding), _tuple2(dilation), groups, _terms_for(input.dtype)).apply(input, weight, bias)


_cache = dict()


def _conv(transpose, weight_shape, stride, padding, output_padding, dilation, groups, terms=3, xg=False, wg=False):
    """xg / wg: the op's input / weight operand is gradient-like (created inside a backward pass): precision 16 keeps those in bf16 (range)."""
    weight_shape = tuple(weight_shape)
    key = (transpose, weight_shape, stride, padding, output_padding, dilation, groups, terms, xg, wg)
    if key in _cache:
        return _cache[key]
    ndim = 2
    kw = dict(stride=stride, padding=padding, dilation=dilation, groups=groups)

    def out_pad_for(input_shape, output_shape):
        # output_padding of the adjoint operator (conv2d_gradfix.py:95-105)
        if transpose:
            return (0, 0)
        return tuple(input_shape[i + 2] - (output_shape[i + 2] - 1) * stride[i] - (1 - 2 * padding[i])
                     - dilation[i] * (weight_shape[i + 2] - 1) for i in range(ndim))

    class Conv2d(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input, weight, bias):
            assert tuple(weight.shape) == weight_shape
            if not transpose:
                out = _primitive_conv(input, weight, bias, stride, padding, dilation, groups, terms, xg, wg)
            else:
                out = _primitive_conv_transpose(input, weight, bias, stride, padding, output_padding, dilation, groups, terms, xg, wg)
            ctx.save_for_backward(input, weight, bias)
            return out

        @staticmethod
        def backward(ctx, grad_output):
            input, weight, bias = ctx.saved_tensors
            gi = gw = gb = None
            if ctx.needs_input_grad[0]:
                p = out_pad_for(input.shape, grad_output.shape)
                gi = _conv(not transpose, weight_shape, stride, padding, p, dilation, groups, terms, True, wg).apply(grad_output, weight, None)
                assert gi.shape == input.shape
            if ctx.needs_input_grad[1] and not weight_gradients_disabled:
                gw = Conv2dGradWeight.apply(grad_output, input, bias)
                assert tuple(gw.shape) == weight_shape
            if ctx.needs_input_grad[2]:
                gb = grad_output.sum([0, 2, 3])
            return gi, gw, gb

    class Conv2dGradWeight(torch.autograd.Function):
        @staticmethod
        def forward(ctx, grad_output, input, bias):
            k = weight_shape[2]
            cin_op = input.shape[1]
            cout_op = grad_output.shape[1]
            use_tc = (tc_enabled and groups == 1 and tuple(dilation) == (1, 1) and weight_shape[2] == weight_shape[3] and k in (1, 3, 5)
                      and input.dtype in (torch.float32, torch.float16) and tc.wgrad_eligible(cin_op, cout_op)
                      and ((not transpose and ((tuple(stride) == (1, 1) and tuple(padding) == (k // 2, k // 2)) or (tuple(stride) == (2, 2) and k == 3 and padding[0] == padding[1])))
                           or (transpose and tuple(stride) == (2, 2) and tuple(padding) == (0, 0) and k == 3)
                           or (transpose and tuple(stride) == (1, 1) and tuple(padding) == (k // 2, k // 2) and tuple(output_padding) == (0, 0))))
            if use_tc:
                tc_stats['tc'] += 1
                if transpose and tuple(stride) == (1, 1):
                    # stride-1 transposed conv == conv with the flipped, transposed kernel and padding k-1-p (= k//2 here): take that conv's
                    # weight gradient and undo the flip / transpose.  This is the weight-gradient of an input-gradient op, i.e. the second-order
                    # term of the R1 penalty (loss.py:238-253).
                    gw = tc.conv_wgrad(grad_output, input, k, 'conv', 1, k // 2, terms, x_is_grad=xg).flip([2, 3]).transpose(0, 1)
                else:
                    gw = tc.conv_wgrad(grad_output, input, k, 'transpose' if transpose else 'conv', stride[0], padding[0], terms, x_is_grad=xg)
                ctx.save_for_backward(grad_output, input)
                return gw
            tc_stats['aten'] += 1
            bias_shape = bias.shape if bias is not None else None
            empty_w = torch.empty(weight_shape, dtype=input.dtype, layout=input.layout, device=input.device)
            gw = torch.ops.aten.convolution_backward(
                grad_output, input, empty_w, bias_sizes=bias_shape, stride=stride, padding=padding, dilation=dilation,
                transposed=transpose, output_padding=output_padding, groups=groups, output_mask=[False, True, False])[1]
            ctx.save_for_backward(grad_output, input)
            return gw

        @staticmethod
        def backward(ctx, g2_gw):
            grad_output, input = ctx.saved_tensors
            g2_go = g2_in = None
            if ctx.needs_input_grad[0]:      # the "weight" operand of these second-order ops is a gradient (g2_gw): keep it in bf16
                g2_go = _conv(transpose, weight_shape, stride, padding, output_padding, dilation, groups, terms, xg, True).apply(input, g2_gw, None)
            if ctx.needs_input_grad[1]:
                p = out_pad_for(input.shape, grad_output.shape)
                g2_in = _conv(not transpose, weight_shape, stride, padding, p, dilation, groups, terms, True, True).apply(grad_output, g2_gw, None)
            return g2_go, g2_in, None

    _cache[key] = Conv2d
    return Conv2d
