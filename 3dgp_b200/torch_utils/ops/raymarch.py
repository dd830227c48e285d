"""Fused tri-plane ray-march op (no reference plugin counterpart: the reference spends ~25 torch ops and several GB of
per-sample tensors on this, tri_plane_renderer.py:126-170).  One forward kernel, one backward kernel
(csrc/raymarch_fwd.cu, raymarch_bwd.cu); nothing per-sample is stored between them.

    rgb, depth, wsum, tfinal = render_rays(planes, w1, b1, w2, b2, ray_o, ray_d, num_steps=48, ...)

planes : [B, 3, C, P, P] (any strides; the kernels want channel-minor storage, i.e. the channels-last layout the
         tri-plane decoder of this package emits -- a strided copy is made otherwise).
Noise  : pass `u_coarse`, `u_fine` ([B, R, N] uniforms) and `sn_coarse`, `sn_fine` (std-normals) to inject the variates
         (parity mode, SURVEY.md 7 "RNG parity"); leave them None for the in-kernel Philox stream (`seed`, `offset`).
"""
import ctypes

import torch

from ... import _lib


TIMING = None   # set to a list to collect (start_event, end_event) pairs around the forward kernel (bench.py roofline)


def planes_channel_minor(planes):
    """[B,3,C,P,P] -> same logical tensor stored as [B, P, P, 3*C] (channel-minor).  No copy if already so."""
    B, K, C, P, P2 = planes.shape
    assert K == 3 and P == P2
    if planes.stride(2) == 1 and planes.stride(1) == C and planes.stride(4) % 4 == 0:
        return planes
    flat = planes.reshape(B, 3 * C, P, P).contiguous(memory_format=torch.channels_last)
    return flat.view(B, 3, C, P, P)


def _opts(B, R, N, P, C, H, o):
    return _lib.RaymarchOpts(
        B=B, R=R, N=N, P=P, C=C, H=H, ray_start=float(o['ray_start']), ray_end=float(o['ray_end']),
        box_half=float(o['box_half']), noise_std=float(o['noise_std']), use_inf_depth=int(o['use_inf_depth']),
        last_back=int(o['last_back']), white_back_end_idx=int(o['white_back_end_idx']),
        clamp_mode={'softplus': 0, 'relu': 1}[o['clamp_mode']], mlp_mode=int(o['mlp_mode']),
        seed=int(o['seed']), offset=int(o['offset']))


def _f32c(t, name):
    if t is None:
        return None
    _lib.require_cuda(t, name)
    return t.detach().to(torch.float32).contiguous()


class _RayMarch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, planes, w1, b1, w2, b2, ray_o, ray_d, u_coarse, u_fine, sn_coarse, sn_fine, o):
        L = _lib.lib()
        _lib.require_cuda(planes, 'planes')
        if planes.dtype not in (torch.float32, torch.float16):
            raise RuntimeError('planes must be float32 or float16')
        pl = planes_channel_minor(planes.detach())
        B, _, C, P, _ = pl.shape
        R = ray_o.shape[1]
        N = int(o['num_steps'])
        H = w1.shape[0]
        ro, rd = _f32c(ray_o, 'ray_o'), _f32c(ray_d, 'ray_d')
        w1c, b1c, w2c, b2c = _f32c(w1, 'w1'), _f32c(b1, 'b1'), _f32c(w2, 'w2'), _f32c(b2, 'b2')
        uc, uf, sc, sf = _f32c(u_coarse, 'u_coarse'), _f32c(u_fine, 'u_fine'), _f32c(sn_coarse, 'sn_coarse'), _f32c(sn_fine, 'sn_fine')
        for t, nm in ((uc, 'u_coarse'), (uf, 'u_fine'), (sc, 'sn_coarse'), (sf, 'sn_fine')):
            if t is not None and t.numel() != B * R * N:
                raise RuntimeError(f'{nm} must have B*R*N = {B * R * N} elements, got {tuple(t.shape)}')
        if tuple(w1c.shape) != (H, C) or tuple(w2c.shape) != (4, H) or b1c.numel() != H or b2c.numel() != 4:
            raise RuntimeError('tri-plane MLP must be 2 layers: [H,C],[H],[4,H],[4]')
        rgb = torch.empty([B, R, 3], dtype=torch.float32, device=pl.device)
        depth = torch.empty([B, R, 1], dtype=torch.float32, device=pl.device)
        wsum = torch.empty([B, R, 1], dtype=torch.float32, device=pl.device)
        tfin = torch.empty([B, R], dtype=torch.float32, device=pl.device)
        opts = _opts(B, R, N, P, C, H, o)
        ev = None
        if TIMING is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        with torch.cuda.device(pl.device):
            rc = L.gp3d_raymarch_forward(
                pl.data_ptr(), _lib.dtype_code(pl), pl.stride(0), pl.stride(1), pl.stride(2), pl.stride(3), pl.stride(4),
                ro.data_ptr(), rd.data_ptr(), w1c.data_ptr(), b1c.data_ptr(), w2c.data_ptr(), b2c.data_ptr(),
                _lib.ptr(uc), _lib.ptr(uf), _lib.ptr(sc), _lib.ptr(sf),
                rgb.data_ptr(), depth.data_ptr(), wsum.data_ptr(), tfin.data_ptr(), ctypes.byref(opts), _lib.stream_ptr())
        if ev is not None:
            ev[1].record()
            TIMING.append(ev + (B, R, N, P, C, pl.element_size()))
        _lib.check(rc, 'raymarch_forward')
        ctx.save_for_backward(pl, w1c, b1c, w2c, b2c, ro, rd,
                              *(t if t is not None else torch.empty(0, device=pl.device) for t in (uc, uf, sc, sf)))
        ctx.o = dict(o)
        ctx.dims = (B, R, N, P, C, H)
        ctx.mark_non_differentiable(wsum, tfin)
        return rgb, depth, wsum, tfin

    @staticmethod
    def backward(ctx, g_rgb, g_depth, _g_wsum, _g_tfin):
        L = _lib.lib()
        pl, w1c, b1c, w2c, b2c, ro, rd, uc, uf, sc, sf = ctx.saved_tensors
        B, R, N, P, C, H = ctx.dims
        nz = lambda t: t if t.numel() else None
        uc, uf, sc, sf = nz(uc), nz(uf), nz(sc), nz(sf)
        g_rgb = torch.zeros([B, R, 3], device=pl.device) if g_rgb is None else g_rgb.to(torch.float32).contiguous()
        g_depth = torch.zeros([B, R, 1], device=pl.device) if g_depth is None else g_depth.to(torch.float32).contiguous()
        g_pl = torch.zeros_like(pl, dtype=torch.float32)   # preserves the channel-minor strides
        assert g_pl.stride() == pl.stride()
        g_w1 = torch.zeros_like(w1c); g_b1 = torch.zeros_like(b1c); g_w2 = torch.zeros_like(w2c); g_b2 = torch.zeros_like(b2c)
        need_rays = ctx.needs_input_grad[5] or ctx.needs_input_grad[6]
        g_ro = torch.empty_like(ro) if need_rays else None
        g_rd = torch.empty_like(rd) if need_rays else None
        opts = _opts(B, R, N, P, C, H, ctx.o)
        with torch.cuda.device(pl.device):
            rc = L.gp3d_raymarch_backward(
                pl.data_ptr(), _lib.dtype_code(pl), pl.stride(0), pl.stride(1), pl.stride(2), pl.stride(3), pl.stride(4),
                ro.data_ptr(), rd.data_ptr(), w1c.data_ptr(), b1c.data_ptr(), w2c.data_ptr(), b2c.data_ptr(),
                _lib.ptr(uc), _lib.ptr(uf), _lib.ptr(sc), _lib.ptr(sf),
                g_rgb.data_ptr(), g_depth.data_ptr(), g_pl.data_ptr(), g_w1.data_ptr(), g_b1.data_ptr(), g_w2.data_ptr(),
                g_b2.data_ptr(), _lib.ptr(g_ro), _lib.ptr(g_rd), ctypes.byref(opts), _lib.stream_ptr())
        _lib.check(rc, 'raymarch_backward')
        g_planes = g_pl if ctx.needs_input_grad[0] else None
        return (g_planes, g_w1, g_b1, g_w2, g_b2, g_ro if ctx.needs_input_grad[5] else None,
                g_rd if ctx.needs_input_grad[6] else None, None, None, None, None, None)


def render_rays(planes, w1, b1, w2, b2, ray_o, ray_d, *, num_steps, ray_start, ray_end, box_size,
                u_coarse=None, u_fine=None, sn_coarse=None, sn_fine=None, density_noise=0.0, use_inf_depth=True,
                last_back=False, white_back_end_idx=0, clamp_mode='softplus', mlp_mode=0, seed=0, offset=0):
    """Returns (rgb [B,R,3], depth [B,R,1], weights_sum [B,R,1], final_transmittance [B,R]) -- the 4-tuple of
    ImportanceRenderer.forward (tri_plane_renderer.py:170).  Differentiable w.r.t. planes, MLP parameters and rays."""
    o = dict(num_steps=num_steps, ray_start=ray_start, ray_end=ray_end, box_half=box_size / 2, noise_std=density_noise,
             use_inf_depth=use_inf_depth, last_back=last_back, white_back_end_idx=white_back_end_idx, clamp_mode=clamp_mode,
             mlp_mode=mlp_mode, seed=seed, offset=offset)
    return _RayMarch.apply(planes, w1, b1, w2, b2, ray_o, ray_d, u_coarse, u_fine, sn_coarse, sn_fine, o)


def generate_rays(c2w, fov, resolution, patch_scales=None, patch_offsets=None):
    """sample_rays (tri_plane_renderer.py:487-527) as one kernel: c2w [B,4,4], fov [B] degrees -> ray_o, ray_d [B, h*w, 3]."""
    L = _lib.lib()
    h, w = resolution
    B = c2w.shape[0]
    c2 = _f32c(c2w, 'c2w'); fv = _f32c(fov, 'fov'); ps = _f32c(patch_scales, 'patch_scales'); po = _f32c(patch_offsets, 'patch_offsets')
    ro = torch.empty([B, h * w, 3], dtype=torch.float32, device=c2.device); rd = torch.empty_like(ro)
    with torch.cuda.device(c2.device):
        rc = L.gp3d_generate_rays(c2.data_ptr(), fv.data_ptr(), _lib.ptr(ps), _lib.ptr(po), B, h, w, ro.data_ptr(), rd.data_ptr(), _lib.stream_ptr())
    _lib.check(rc, 'generate_rays')
    return ro, rd


def _rays_from_camera_torch(c2w, fov, h, w, ps, po):
    """Differentiable restatement of the in-kernel ray generator (only the backward of a TRAINED camera uses it: d(ray) -> d(cam2world, fov))."""
    dev = c2w.device
    B = c2w.shape[0]
    x, y = torch.meshgrid(torch.linspace(-1, 1, w, device=dev), torch.linspace(1, -1, h, device=dev), indexing='ij')
    x = x.T.flatten().unsqueeze(0).repeat(B, 1); y = y.T.flatten().unsqueeze(0).repeat(B, 1)
    if ps is not None:
        x = (x + 1.0) * ps[:, 0].view(B, 1) - 1.0 + po[:, 0].view(B, 1) * 2.0
        y = (y + 1.0) * ps[:, 1].view(B, 1) - 1.0 + po[:, 1].view(B, 1) * 2.0
    z = (-1.0 / torch.tan(fov.view(B, 1) / 360 * 2 * 3.141592653589793 * 0.5)).expand(B, h * w)
    d = torch.stack([x, y, z], dim=2)
    d = d / torch.norm(d, dim=2, keepdim=True)
    rd = torch.bmm(c2w[:, :3, :3], d.permute(0, 2, 1)).permute(0, 2, 1)
    ro = c2w[:, :3, 3].unsqueeze(1).expand(B, h * w, 3)
    return ro, rd


class _RayMarchCam(torch.autograd.Function):
    """Same render with the rays generated inside the forward kernel from (cam2world, fov, patch transform): 4 x 4 pixel tiles, no ray tensors in HBM."""

    @staticmethod
    def forward(ctx, planes, w1, b1, w2, b2, c2w, fov, patch_scales, patch_offsets, u_coarse, u_fine, sn_coarse, sn_fine, o):
        L = _lib.lib()
        _lib.require_cuda(planes, 'planes')
        if planes.dtype not in (torch.float32, torch.float16):
            raise RuntimeError('planes must be float32 or float16')
        pl = planes_channel_minor(planes.detach())
        B, _, C, P, _ = pl.shape
        h, w = o['resolution']
        R = h * w
        N = int(o['num_steps'])
        H = w1.shape[0]
        c2, fv = _f32c(c2w, 'c2w'), _f32c(fov, 'fov')
        ps, po = _f32c(patch_scales, 'patch_scales'), _f32c(patch_offsets, 'patch_offsets')
        if tuple(c2.shape) != (B, 4, 4) or fv.numel() != B:
            raise RuntimeError('camera must be cam2world [B,4,4] and fov [B]')
        w1c, b1c, w2c, b2c = _f32c(w1, 'w1'), _f32c(b1, 'b1'), _f32c(w2, 'w2'), _f32c(b2, 'b2')
        uc, uf, sc, sf = _f32c(u_coarse, 'u_coarse'), _f32c(u_fine, 'u_fine'), _f32c(sn_coarse, 'sn_coarse'), _f32c(sn_fine, 'sn_fine')
        for t, nm in ((uc, 'u_coarse'), (uf, 'u_fine'), (sc, 'sn_coarse'), (sf, 'sn_fine')):
            if t is not None and t.numel() != B * R * N:
                raise RuntimeError(f'{nm} must have B*R*N = {B * R * N} elements, got {tuple(t.shape)}')
        if tuple(w1c.shape) != (H, C) or tuple(w2c.shape) != (4, H) or b1c.numel() != H or b2c.numel() != 4:
            raise RuntimeError('tri-plane MLP must be 2 layers: [H,C],[H],[4,H],[4]')
        rgb = torch.empty([B, R, 3], dtype=torch.float32, device=pl.device)
        depth = torch.empty([B, R, 1], dtype=torch.float32, device=pl.device)
        wsum = torch.empty([B, R, 1], dtype=torch.float32, device=pl.device)
        tfin = torch.empty([B, R], dtype=torch.float32, device=pl.device)
        opts = _opts(B, R, N, P, C, H, o)
        cam = _lib.RaymarchCam(c2.data_ptr(), fv.data_ptr(), _lib.ptr(ps), _lib.ptr(po), h, w)
        ev = None
        if TIMING is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        with torch.cuda.device(pl.device):
            rc = L.gp3d_raymarch_forward_cam(
                pl.data_ptr(), _lib.dtype_code(pl), pl.stride(0), pl.stride(1), pl.stride(2), pl.stride(3), pl.stride(4), ctypes.byref(cam),
                w1c.data_ptr(), b1c.data_ptr(), w2c.data_ptr(), b2c.data_ptr(), _lib.ptr(uc), _lib.ptr(uf), _lib.ptr(sc), _lib.ptr(sf),
                rgb.data_ptr(), depth.data_ptr(), wsum.data_ptr(), tfin.data_ptr(), ctypes.byref(opts), _lib.stream_ptr())
        if ev is not None:
            ev[1].record()
            TIMING.append(ev + (B, R, N, P, C, pl.element_size()))
        _lib.check(rc, 'raymarch_forward_cam')
        e = lambda t: t if t is not None else torch.empty(0, device=pl.device)
        ctx.save_for_backward(pl, w1c, b1c, w2c, b2c, c2, fv, e(ps), e(po), e(uc), e(uf), e(sc), e(sf))
        ctx.o = dict(o)
        ctx.dims = (B, R, N, P, C, H, h, w)
        ctx.mark_non_differentiable(wsum, tfin)
        return rgb, depth, wsum, tfin

    @staticmethod
    def backward(ctx, g_rgb, g_depth, _g_wsum, _g_tfin):
        L = _lib.lib()
        pl, w1c, b1c, w2c, b2c, c2, fv, ps, po, uc, uf, sc, sf = ctx.saved_tensors
        B, R, N, P, C, H, h, w = ctx.dims
        nz = lambda t: t if t.numel() else None
        ps, po, uc, uf, sc, sf = nz(ps), nz(po), nz(uc), nz(uf), nz(sc), nz(sf)
        ro, rd = generate_rays(c2, fv, (h, w), ps, po)
        g_rgb = torch.zeros([B, R, 3], device=pl.device) if g_rgb is None else g_rgb.to(torch.float32).contiguous()
        g_depth = torch.zeros([B, R, 1], device=pl.device) if g_depth is None else g_depth.to(torch.float32).contiguous()
        g_pl = torch.zeros_like(pl, dtype=torch.float32)
        assert g_pl.stride() == pl.stride()
        g_w1 = torch.zeros_like(w1c); g_b1 = torch.zeros_like(b1c); g_w2 = torch.zeros_like(w2c); g_b2 = torch.zeros_like(b2c)
        need_cam = ctx.needs_input_grad[5] or ctx.needs_input_grad[6]
        g_ro = torch.empty_like(ro) if need_cam else None
        g_rd = torch.empty_like(rd) if need_cam else None
        opts = _opts(B, R, N, P, C, H, ctx.o)
        with torch.cuda.device(pl.device):
            rc = L.gp3d_raymarch_backward(
                pl.data_ptr(), _lib.dtype_code(pl), pl.stride(0), pl.stride(1), pl.stride(2), pl.stride(3), pl.stride(4),
                ro.data_ptr(), rd.data_ptr(), w1c.data_ptr(), b1c.data_ptr(), w2c.data_ptr(), b2c.data_ptr(),
                _lib.ptr(uc), _lib.ptr(uf), _lib.ptr(sc), _lib.ptr(sf),
                g_rgb.data_ptr(), g_depth.data_ptr(), g_pl.data_ptr(), g_w1.data_ptr(), g_b1.data_ptr(), g_w2.data_ptr(),
                g_b2.data_ptr(), _lib.ptr(g_ro), _lib.ptr(g_rd), ctypes.byref(opts), _lib.stream_ptr())
        _lib.check(rc, 'raymarch_backward')
        g_c2w = g_fov = None
        if need_cam:      # chain d(ray_o), d(ray_d) -> d(cam2world), d(fov) through the ray generator (tiny tensors)
            with torch.enable_grad():
                c2g = c2.detach().requires_grad_(True); fvg = fv.detach().requires_grad_(True)
                ro_t, rd_t = _rays_from_camera_torch(c2g, fvg, h, w, ps, po)
                g_c2w, g_fov = torch.autograd.grad([ro_t, rd_t], [c2g, fvg], [g_ro, g_rd])
            if not ctx.needs_input_grad[5]:
                g_c2w = None
            if not ctx.needs_input_grad[6]:
                g_fov = None
        g_planes = g_pl if ctx.needs_input_grad[0] else None
        return (g_planes, g_w1, g_b1, g_w2, g_b2, g_c2w, g_fov, None, None, None, None, None, None, None)


def render_camera(planes, w1, b1, w2, b2, c2w, fov, resolution, patch_scales=None, patch_offsets=None, *, num_steps, ray_start, ray_end, box_size,
                  u_coarse=None, u_fine=None, sn_coarse=None, sn_fine=None, density_noise=0.0, use_inf_depth=True,
                  last_back=False, white_back_end_idx=0, clamp_mode='softplus', mlp_mode=2, seed=0, offset=0):
    """render_rays with the rays of a pinhole camera per image generated inside the kernel (sample_rays + ImportanceRenderer.forward in ONE launch):
    c2w [B,4,4] cam2world, fov [B] degrees, resolution (h, w), optional patch transform.  Differentiable w.r.t. planes, MLP parameters, c2w and fov."""
    o = dict(num_steps=num_steps, ray_start=ray_start, ray_end=ray_end, box_half=box_size / 2, noise_std=density_noise,
             use_inf_depth=use_inf_depth, last_back=last_back, white_back_end_idx=white_back_end_idx, clamp_mode=clamp_mode,
             mlp_mode=mlp_mode, seed=seed, offset=offset, resolution=(int(resolution[0]), int(resolution[1])))
    return _RayMarchCam.apply(planes, w1, b1, w2, b2, c2w, fov, patch_scales, patch_offsets, u_coarse, u_fine, sn_coarse, sn_fine, o)
