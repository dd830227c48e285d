"""Fused modulated-convolution layer for the tri-plane decoder's training path (fp32, CUDA):

    y = lrelu( FIR_up?( conv( x * styles , W ) ) * dcoefs + noise * strength + bias ) * gain

i.e. networks_stylegan2.py:67-76 + :142-144 (x*styles -> conv2d_resample -> fma -> bias_act) as ONE autograd node whose
forward is   split(x*s -> bf16 hi/lo)  ->  tcgen05 conv (stride-1 or polyphase up-2)  [-> upfirdn2d]  ->  demod+noise+bias+act
and whose backward is   demod_act_bwd (dt, dc, g_bias, g_dcoefs, g_strength in one pass)  [-> upfirdn2d adjoint]  ->  split  ->
tcgen05 input-gradient conv + tcgen05 weight-gradient GEMM  ->  modulate_bwd (dx, g_styles in one pass).
Nothing but x, y and the bf16 operand pair is kept for backward; no x*styles / pre-activation / FIR intermediates survive.
First-order only (the reference differentiates G twice only when pl_weight > 0, which the 3dgp config sets to 0).
"""
import ctypes
import os

import torch

from ... import _lib
from . import tc, upfirdn2d


def eligible(x, weight, up, conv_clamp):
    """Forward (Cin -> Cout), input-gradient (Cout padded to whole 64-channel blocks -> Cin) and weight-gradient shapes must all be covered by the
    tensor-core kernels; anything else (e.g. Cin = 192 from a non-power-of-two cbase) takes the unfused composition."""
    Cout, Cin, k, _ = weight.shape
    Cp = ((Cout + 63) // 64) * 64
    return (x.is_cuda and x.dtype == torch.float32 and conv_clamp is None and up in (1, 2) and k in (1, 3) and (up == 1 or k == 3)
            and tc.channels_eligible(Cin, Cout) and tc.channels_eligible(Cp, Cin) and tc.wgrad_eligible(Cin, Cp) and Cin % 4 == 0 and Cout % 4 == 0)


def _fir_tma_ok(C, fir):
    return C % 32 == 0 and fir is not None and fir.dtype == torch.float32 and tuple(fir.shape) == (4, 4) and fir.is_contiguous()


def _nhwc(t):
    return t.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)


G_TERMS = int(os.environ.get('GP3D_G_TERMS', '3'))     # precision of the tri-plane decoder's forward / input-gradient convolutions: 3 = bf16x3, 2 = x2w16 (ops.tc.operand_formats)


def _conv_fwd(xh, xl, wh, wl, N, H, W, Cin, Cout, k, up):
    """Raw convolution of the up-sampling layers: stride-2 transposed conv as four polyphase tap convolutions (phases of one launch) -> [N, 2H+1, 2W+1, Cout]."""
    dev = xh.device
    assert up == 2
    Ho, Wo = 2 * H + 1, 2 * W + 1
    y = torch.empty([N, Ho, Wo, Cout], dtype=torch.float32, device=dev)
    tc.conv_transpose_s2_launch(xh, xl, wh, wl, y, N, H, W, Cin, Cout)
    return y


class _ModConvLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, styles, dcoefs, noise, noise_strength, bias, up, fir, act, alpha, gain, terms):
        L = _lib.lib()
        upfirdn2d._init()
        tc.stats['fused'] += 1
        N, Cin, H, W = x.shape
        Cout, _, k, _ = weight.shape
        dev = x.device
        xn = _nhwc(x)
        st = styles.to(torch.float32).contiguous()
        xh, xl = tc.split_bf16(xn, styles=st, fp16=(terms == 2))          # terms 2: fp16 pair x fp16 weight (operands of one MMA share their format)
        wh, wl = tc.weight_operands(weight, 'fwd', lambda w_: w_.permute(0, 2, 3, 1), terms)
        d = dcoefs.to(torch.float32).contiguous() if dcoefs is not None else None
        nz = None
        nps = 0
        if noise is not None:
            nz = (noise.to(torch.float32) * noise_strength).contiguous()          # [1|N, 1, Ho, Wo] or [Ho, Wo]
            nps = 1 if (nz.dim() == 4 and nz.shape[0] == N and N > 1) else 0
        b = bias.to(torch.float32).contiguous() if bias is not None else None
        if up == 1:   # demodulation, noise, bias and activation ride in the conv kernel's TMEM -> HBM epilogue: the raw conv output is never stored
            Ho, Wo = H, W
            y = torch.empty([N, H, W, Cout], dtype=torch.float32, device=dev)
            epi = _lib.ConvEpilogue(_lib.ptr(d), _lib.ptr(nz), _lib.ptr(b), nps, 3 if act == 'lrelu' else 1, float(alpha), float(gain), -1.0)
            tc.conv_launch(xh, xl, wh, wl, y, N, H, W, Cin, Cout, k * k, tc.same_taps(k), 1, H, W, H, W, 1, 1, 0, 0, epi=epi, what='conv2d_nhwc_act')
        else:
            c = _conv_fwd(xh, xl, wh, wl, N, H, W, Cin, Cout, k, up)
            # FIR after the stride-2 transposed conv: pad 1, gain up^2 (conv2d_resample.py:119-126 for k=3, fw=4)
            Ho, Wo = 2 * H, 2 * W
            y = torch.empty([N, Ho, Wo, Cout], dtype=torch.float32, device=dev)
            if _fir_tma_ok(Cout, fir):    # TMA-staged FIR with demodulation / noise / bias / activation in its epilogue: the filtered tensor is never stored
                epi = _lib.ConvEpilogue(_lib.ptr(d), _lib.ptr(nz), _lib.ptr(b), nps, 3 if act == 'lrelu' else 1, float(alpha), float(gain))
                with torch.cuda.device(dev):
                    rc = L.gp3d_fir4_nhwc(c.data_ptr(), fir.data_ptr(), 0, 4.0, N, 2 * H + 1, 2 * W + 1, Cout, 1, 1, 1, 1, y.data_ptr(), None, None,
                                          ctypes.byref(epi), _lib.stream_ptr())
                _lib.check(rc, 'fir4_nhwc')
            else:
                c = upfirdn2d._plugin.upfirdn2d(c.permute(0, 3, 1, 2), fir, 1, 1, 1, 1, 1, 1, 1, 1, False, 4.0).permute(0, 2, 3, 1)
                with torch.cuda.device(dev):
                    rc = L.gp3d_demod_act(c.data_ptr(), _lib.ptr(d), _lib.ptr(nz), nps, _lib.ptr(b), y.data_ptr(), 0, N, Cout, Ho * Wo, 1,
                                          3 if act == 'lrelu' else 1, float(alpha), float(gain), -1.0, _lib.stream_ptr())
                _lib.check(rc, 'demod_act')
        if terms == 2:      # the gradient products run as bf16x3 (range): the backward re-splits the modulated input into a bf16 pair
            xh = xl = torch.empty(0, device=dev)
        ctx.save_for_backward(xn, xh, xl, weight, st, d if d is not None else torch.empty(0, device=dev), y,
                              noise if noise is not None else torch.empty(0, device=dev),
                              noise_strength if noise is not None else torch.empty(0, device=dev),
                              b if b is not None else torch.empty(0, device=dev), fir if fir is not None else torch.empty(0, device=dev))
        ctx.cfg = (N, Cin, H, W, Cout, k, up, act, float(alpha), float(gain), nps, Ho, Wo, terms)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        from . import conv2d_gradfix
        L = _lib.lib()
        xn, xh, xl, weight, st, d, y, noise, noise_strength, b, fir = ctx.saved_tensors
        need_dx = ctx.needs_input_grad[0] or ctx.needs_input_grad[2]
        need_gw = ctx.needs_input_grad[1] and not conv2d_gradfix.weight_gradients_disabled
        N, Cin, H, W, Cout, k, up, act, alpha, gain, nps, Ho, Wo, terms = ctx.cfg
        dev = dy.device
        if terms == 2:
            terms = 3
            if need_gw:
                xh, xl = tc.split_bf16(xn, styles=st)
        has_d, has_n, has_b = d.numel() > 0, noise.numel() > 0, b.numel() > 0
        dyn = _nhwc(dy.to(torch.float32))
        # channels of the gradient operand are zero-padded to whole 64-channel TMA blocks (toRGB: 96 -> 128)
        Cp = ((Cout + 63) // 64) * 64
        if up == 1:    # no FIR in between: the activation backward writes the conv kernels' bf16 (hi, lo) operands directly
            dc = None
            dch = torch.empty([N, Ho, Wo, Cp], dtype=torch.bfloat16, device=dev)
            dcl = torch.empty_like(dch)
        else:
            dc = torch.empty_like(dyn)
            dch = dcl = None
        g_d = torch.zeros_like(d) if has_d else None
        g_b = torch.zeros([Cout], dtype=torch.float32, device=dev) if has_b else None
        nz_img = noise.to(torch.float32).contiguous() if has_n else None                 # unscaled noise image
        ns = noise_strength.to(torch.float32).reshape(1).contiguous() if has_n else None
        g_ns = torch.zeros([1], dtype=torch.float32, device=dev) if has_n else None
        with torch.cuda.device(dev):
            rc = L.gp3d_demod_act_bwd_split(dyn.data_ptr(), y.data_ptr(), _lib.ptr(d if has_d else None), _lib.ptr(nz_img), _lib.ptr(ns), nps,
                                            _lib.ptr(b if has_b else None), _lib.ptr(dc), _lib.ptr(dch), _lib.ptr(dcl), Cp, _lib.ptr(g_d), _lib.ptr(g_b), _lib.ptr(g_ns),
                                            N, Ho * Wo, Cout, 3 if act == 'lrelu' else 1, alpha, gain, _lib.stream_ptr())
        _lib.check(rc, 'demod_act_bwd')
        if g_ns is not None:
            g_ns = g_ns.reshape(noise_strength.shape)
        if up == 2:   # adjoint of the FIR (upfirdn2d.py:250-269): same filter, flipped, padding p = fw - pad - 1 = 2
            if _fir_tma_ok(Cout, fir) and Cp == Cout:   # TMA-staged adjoint FIR that writes the bf16 (hi, lo) operand pair directly
                dch = torch.empty([N, Ho + 1, Wo + 1, Cp], dtype=torch.bfloat16, device=dev)
                dcl = torch.empty_like(dch)
                with torch.cuda.device(dev):
                    rc = L.gp3d_fir4_nhwc(dc.data_ptr(), fir.data_ptr(), 1, 4.0, N, Ho, Wo, Cout, 2, 2, 2, 2, None, dch.data_ptr(), dcl.data_ptr(),
                                          None, _lib.stream_ptr())
                _lib.check(rc, 'fir4_nhwc')
            else:
                dc = upfirdn2d._plugin.upfirdn2d(dc.permute(0, 3, 1, 2), fir, 1, 1, 1, 1, 2, 2, 2, 2, True, 4.0).permute(0, 2, 3, 1).contiguous()
                dch, dcl = tc.split_bf16(dc, pad_to=Cp)
        # input gradient
        dxs = torch.empty([N, H, W, Cin], dtype=torch.float32, device=dev) if need_dx else None
        assert tc.channels_eligible(Cp, Cin)
        if not need_dx:
            pass
        elif up == 1:
            wdh, wdl = tc.weight_operands(weight, 'dgrad1', lambda w_: w_.flip([2, 3]).permute(1, 2, 3, 0), terms, pad_to=Cp)    # [Cin,k,k,Cout(+pad)]
            tc.conv_launch(dch, dcl, wdh, wdl, dxs, N, H, W, Cp, Cin, k * k, tc.same_taps(k), 1, H, W, H, W, 1, 1, 0, 0, what='conv2d_nhwc (input gradient)')
        else:   # dx[i,j] = sum dc1[2i+ky, 2j+kx] w[co][ci][ky][kx]: strided gather over the (2H+1)^2 gradient
            wdh, wdl = tc.weight_operands(weight, 'dgrad2', lambda w_: w_.permute(1, 2, 3, 0), terms, pad_to=Cp)          # [Cin,ky,kx,Cout]
            taps = [(ky, kx, ky * 3 + kx) for ky in range(3) for kx in range(3)]
            tc._taps_launch(dch, dcl, wdh, wdl, dxs, N, 2 * H + 1, 2 * W + 1, Cp, Cin, 9, taps, 2, H, W, H, W, 1, 1, 0, 0)
        # weight gradient
        gw = None
        if not need_gw:
            pass
        elif tc.wgrad_eligible(Cin, Cp):
            if up == 1:
                taps = [(0, 0, ky - k // 2, kx - k // 2, ky * k + kx) for ky in range(k) for kx in range(k)]
                sa, sb, HoP, WoP, Hd, Wd = 1, 1, H, W, H, W
            else:
                taps = [(ky, kx, 0, 0, ky * 3 + kx) for ky in range(3) for kx in range(3)]
                sa, sb, HoP, WoP, Hd, Wd = 2, 1, H, W, 2 * H + 1, 2 * W + 1
            gw = torch.zeros([Cp, k * k, Cin], dtype=torch.float32, device=dev)
            tc.wgrad_launch(dch, dcl, xh, xl, gw, N, Hd, Wd, Cp, H, W, Cin, k * k, taps, sa, sb, HoP, WoP)
            gw = gw[:Cout].view(Cout, k, k, Cin).permute(0, 3, 1, 2).to(weight.dtype)
        else:
            raise RuntimeError('modconv: weight-gradient shape not covered by the tensor-core kernel (Cin %% 64 != 0)')
        # through the modulation
        dx = g_s = None
        if need_dx:
            dx = torch.empty_like(dxs)
            g_s = torch.zeros_like(st)
            with torch.cuda.device(dev):
                rc = L.gp3d_modulate_bwd(dxs.data_ptr(), xn.data_ptr(), st.data_ptr(), dx.data_ptr(), g_s.data_ptr(), N, H * W, Cin, _lib.stream_ptr())
            _lib.check(rc, 'modulate_bwd')
            dx = dx.permute(0, 3, 1, 2)
        return (dx, gw, g_s, g_d, None, g_ns, g_b, None, None, None, None, None, None)


def modconv_layer(x, weight, styles, dcoefs=None, noise=None, noise_strength=None, bias=None, up=1, fir=None, act='lrelu', alpha=0.2, gain=1.0, terms=None):
    if noise is not None and not torch.is_tensor(noise_strength):
        noise_strength = torch.as_tensor(float(noise_strength if noise_strength is not None else 1.0), device=x.device)
    return _ModConvLayer.apply(x, weight, styles, dcoefs, noise, noise_strength, bias, up, fir, act, alpha, gain, G_TERMS if terms is None else terms)


# ---------------------------------------------------------------------------------------------------------------------------------
# Conv2dLayer of the discriminator / depth adaptor (layers.py:228-241) as one first-order autograd node:
#   y = clamp( act( conv( x * s[n, ci]?, w * wgain ) + b ) * gain )
# forward : split (hyper-modulation fused) -> tcgen05 conv whose epilogue applies bias / activation / gain / clamp
# backward: act backward (clamp mask, bias gradient) emitting the bf16 operand -> tcgen05 input-gradient conv (+ modulate_bwd) + weight gradient.
# terms = 3 (bf16x3), 16 (fp16 x fp16 forward, bf16 x bf16 gradient products: the blocks the reference runs in fp16) or 1 (bf16 throughout).
# Stride-1 'same' shapes only.

def conv_act_eligible(x, weight, k, up, down, padding, act, cin, cout):
    return (x.is_cuda and x.dtype == torch.float32 and up == 1 and down == 1 and k in (1, 3) and padding == k // 2 and act in ('linear', 'lrelu')
            and cin % 64 == 0 and cout % 64 == 0 and tc.channels_eligible(cin, cout) and tc.channels_eligible(cout, cin))


class _ConvBiasAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, s, wgain, act, alpha, gain, clamp, terms):
        L = _lib.lib()
        tc.stats['fused'] += 1
        N, Cin, H, W = x.shape
        Cout, _, k, _ = weight.shape
        dev = x.device
        xn = _nhwc(x)
        st = s.to(torch.float32).contiguous() if s is not None else None
        x_lo, _, x_fp16, _ = tc.operand_formats(terms)
        xh, xl = tc.split_bf16(xn, styles=st, want_lo=x_lo, fp16=x_fp16)
        wh, wl = tc.weight_operands(weight, ('cfwd', float(wgain)), lambda w_: (w_ * wgain).permute(0, 2, 3, 1), terms)
        b = bias.to(torch.float32).contiguous() if bias is not None else None
        # channel-minor storage behind an ordinary NCHW-shaped tensor (not a view: DiscriminatorBlock adds into the skip output in place)
        y = torch.empty([N, Cout, H, W], dtype=torch.float32, device=dev, memory_format=torch.channels_last)
        epi = _lib.ConvEpilogue(None, None, _lib.ptr(b), 0, 3 if act == 'lrelu' else 1, float(alpha), float(gain), float(clamp) if clamp is not None else -1.0)
        tc.conv_launch(xh, xl, wh, wl, y, N, H, W, Cin, Cout, k * k, tc.same_taps(k), 1, H, W, H, W, 1, 1, 0, 0, epi=epi, what='conv2d_nhwc_act')
        # linear, unclamped layers (the residual skip) do not need their output in the backward: callers may update it in place (y.add_(x))
        keep_y = (act == 'lrelu') or (clamp is not None)
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(xh, xl if xl is not None else empty, weight, st if st is not None else empty, xn if st is not None else empty,
                              y if keep_y else empty)
        ctx.cfg = (N, Cin, H, W, Cout, k, float(wgain), act, float(alpha), float(gain), clamp, terms, bias is not None)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        from . import conv2d_gradfix
        L = _lib.lib()
        xh, xl, weight, st, xn, y = ctx.saved_tensors
        N, Cin, H, W, Cout, k, wgain, act, alpha, gain, clamp, terms, has_b = ctx.cfg
        dev = dy.device
        xl = xl if xl.numel() else None
        gterms = 3 if terms == 3 else 1             # gradient products: bf16x3, or one bf16 x bf16 product (gradients need bf16's range)
        dyn = _nhwc(dy.to(torch.float32))
        dch = torch.empty([N, H, W, Cout], dtype=torch.bfloat16, device=dev)
        dcl = torch.empty_like(dch) if terms == 3 else None
        g_b = torch.zeros([Cout], dtype=torch.float32, device=dev) if (has_b and ctx.needs_input_grad[2]) else None
        yref = y if y.numel() else dyn            # linear / unclamped: the slope does not depend on y
        with torch.cuda.device(dev):
            rc = L.gp3d_act_bwd_split(dyn.data_ptr(), yref.data_ptr(), None, dch.data_ptr(), _lib.ptr(dcl), Cout, _lib.ptr(g_b), N, H * W, Cout,
                                      3 if act == 'lrelu' else 1, alpha, gain, float(clamp) if clamp is not None else -1.0, _lib.stream_ptr())
        _lib.check(rc, 'act_bwd_split')
        dx = g_s = gw = None
        if ctx.needs_input_grad[0] or (st.numel() and ctx.needs_input_grad[3]):
            wdh, wdl = tc.weight_operands(weight, ('cadj', wgain), lambda w_: (w_ * wgain).flip([2, 3]).permute(1, 2, 3, 0), gterms)      # [Cin,k,k,Cout]
            dxs = torch.empty([N, H, W, Cin], dtype=torch.float32, device=dev)
            tc.conv_launch(dch, dcl, wdh, wdl, dxs, N, H, W, Cout, Cin, k * k, tc.same_taps(k), 1, H, W, H, W, 1, 1, 0, 0, what='conv2d_nhwc (input gradient)')
            if st.numel():
                dxo = torch.empty_like(dxs)
                g_s = torch.zeros_like(st)
                with torch.cuda.device(dev):
                    rc = L.gp3d_modulate_bwd(dxs.data_ptr(), xn.data_ptr(), st.data_ptr(), dxo.data_ptr(), g_s.data_ptr(), N, H * W, Cin, _lib.stream_ptr())
                _lib.check(rc, 'modulate_bwd')
                dx = dxo.permute(0, 3, 1, 2)
            else:
                dx = dxs.permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1] and not conv2d_gradfix.weight_gradients_disabled:
            taps = [(0, 0, ky - k // 2, kx - k // 2, ky * k + kx) for ky in range(k) for kx in range(k)]
            gwf = torch.zeros([Cout, k * k, Cin], dtype=torch.float32, device=dev)
            if xh.dtype == torch.float16:           # the forward's fp16 operand -> bf16 for the product with the bf16 gradient (same-format rule)
                xh = xh.to(torch.bfloat16)
            tc.wgrad_launch(dch, dcl, xh, xl, gwf, N, H, W, Cout, H, W, Cin, k * k, taps, 1, 1, H, W)
            gw = (gwf.view(Cout, k, k, Cin).permute(0, 3, 1, 2) * wgain).to(weight.dtype)
        return dx, gw, g_b, g_s, None, None, None, None, None, None


def conv_bias_act(x, weight, bias=None, s=None, wgain=1.0, act='linear', alpha=0.2, gain=1.0, clamp=None, terms=3):
    return _ConvBiasAct.apply(x, weight, bias, s, wgain, act, alpha, gain, clamp, terms)
