"""One launch of every hot kernel at a BASELINE shape (B=8), for `ncu --set full`.  python tools/ncu_kernels.py"""
import ctypes, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
tc = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
_lib = importlib.import_module('3dgp_b200._lib')
up = importlib.import_module('3dgp_b200.torch_utils.ops.upfirdn2d')
ba = importlib.import_module('3dgp_b200.torch_utils.ops.bias_act')
rmod = importlib.import_module('3dgp_b200.torch_utils.ops.raymarch')
stepm = importlib.import_module('3dgp_b200.training.step')
L = _lib.lib()
dev = torch.device('cuda')
torch.manual_seed(0)
B = 8
s = _lib.stream_ptr()
# G layer b256.conv1: 256 -> 256 @ 256^2, 3x3, bf16x3 with the fused epilogue; its weight gradient; its input gradient operands
N, H, C = B, 256, 256
x = torch.randn(N, H, H, C, device=dev); w = torch.randn(C, 3, 3, C, device=dev) / 48
st = torch.rand(N, C, device=dev) + 0.5
xh, xl = tc.split_bf16(x, styles=st)                              # split_bf16_cm_kernel<float>
wh, wl = tc.split_bf16(w)
d = torch.rand(N, C, device=dev) + 0.5; nz = torch.randn(N, H, H, device=dev) * 0.1; b = torch.zeros(C, device=dev)
y = torch.empty(N, H, H, C, device=dev)
epi = _lib.ConvEpilogue(d.data_ptr(), nz.data_ptr(), b.data_ptr(), 1, 3, 0.2, 1.4142135)
_lib.check(L.gp3d_conv2d_nhwc_bf16x3_act(xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr(), y.data_ptr(), N, H, H, C, C, 3, ctypes.byref(epi), s), 'conv_act')
dy = torch.randn_like(y)
dch = torch.empty(N, H, H, C, dtype=torch.bfloat16, device=dev); dcl = torch.empty_like(dch)
g_d = torch.zeros_like(d); g_b = torch.zeros(C, device=dev); g_ns = torch.zeros(1, device=dev); ns = torch.ones(1, device=dev)
_lib.check(L.gp3d_demod_act_bwd_split(dy.data_ptr(), y.data_ptr(), d.data_ptr(), nz.data_ptr(), ns.data_ptr(), 1, b.data_ptr(), None, dch.data_ptr(), dcl.data_ptr(), C,
                                      g_d.data_ptr(), g_b.data_ptr(), g_ns.data_ptr(), N, H * H, C, 3, 0.2, 1.4142135, s), 'demod_act_bwd')
gw = torch.zeros(C, 9, C, device=dev)
taps = [(0, 0, ky - 1, kx - 1, ky * 3 + kx) for ky in range(3) for kx in range(3)]
arr = (ctypes.c_int * 45)(*[v for t in taps for v in t])
_lib.check(L.gp3d_wgrad_taps_nhwc(dch.data_ptr(), dcl.data_ptr(), xh.data_ptr(), xl.data_ptr(), gw.data_ptr(), N, H, H, C, H, H, C, 9, 9, ctypes.cast(arr, ctypes.c_void_p), 1, 1, H, H, s), 'wgrad')
dxs = torch.randn_like(x); dx = torch.empty_like(x); g_s = torch.zeros_like(st)
_lib.check(L.gp3d_modulate_bwd(dxs.data_ptr(), x.data_ptr(), st.data_ptr(), dx.data_ptr(), g_s.data_ptr(), N, H * H, C, s), 'modulate_bwd')
# D layer: 1024 -> 1024 @ 64^2 fp16, single-term bf16, 256-wide tiles
xd = torch.randn(B, 1024, 64, 64, device=dev).half().contiguous(memory_format=torch.channels_last)
wd = (torch.randn(1024, 1024, 3, 3, device=dev) / 96).half()
tc.conv2d_forward(xd, wd, 1)
# FIR after the up-sampling conv, skip-image upsample, D's fp16 FIR, bias_act
f4 = up.setup_filter([1, 3, 3, 1], device=dev)
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
up.upfirdn2d(cl(torch.randn(B, 128, 513, 513, device=dev)), f4, padding=1, gain=4)
up.upsample2d(cl(torch.randn(B, 96, 256, 256, device=dev)), f4)
up.upfirdn2d(xd, f4, padding=2)
up.upsample2d(torch.randn(B, 256, 128, 128, device=dev).half(), f4)          # NCHW fp16 (BASELINE configs[3] shape)
up.downsample2d(torch.randn(B, 256, 256, 256, device=dev).half(), f4)
ba.bias_act(cl(torch.randn(B, 128, 512, 512, device=dev)), torch.zeros(128, device=dev), act='lrelu')
ba.bias_act(xd, torch.zeros(1024, device=dev).half(), act='lrelu', clamp=256)
# optimiser: 64 M parameters
net = torch.nn.Linear(8192, 8192, bias=False).to(dev)
ema = torch.nn.Linear(8192, 8192, bias=False).to(dev).requires_grad_(False)
opt = stepm.FlatAdam(net, lr=2e-3, betas=(0.0, 0.99), ema_module=ema)
opt.zero_grad(); opt.flat_g.normal_(); opt.active.update(range(len(opt.params)))
opt.step(ema_beta=0.99)
# toRGB of the 512^2 block: 128 -> 96 channels, 1x1, bf16x3 with the fused epilogue (store-heavy: 2 k-blocks per 48 KB output tile)
N, H, C, Co = B, 512, 128, 96
x = torch.randn(N, H, H, C, device=dev); w = torch.randn(Co, 1, 1, C, device=dev) / 12
xh, xl = tc.split_bf16(x); wh, wl = tc.split_bf16(w)
y = torch.empty(N, H, H, Co, device=dev)
bb = torch.zeros(Co, device=dev)
epi = _lib.ConvEpilogue(None, None, bb.data_ptr(), 0, 1, 0.2, 1.0)
_lib.check(L.gp3d_conv2d_nhwc_bf16x3_act(xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr(), y.data_ptr(), N, H, H, C, Co, 1, ctypes.byref(epi), s), 'torgb')
# ray-march forward + backward, config 3 at B=8 (rays generated in the kernel)
dn = importlib.import_module('3dgp_b200.dnnlib'); ru = importlib.import_module('3dgp_b200.training.rendering_utils')
inp = bench.raymarch_inputs(B, dev, seed=0)
dd = {k: (v.to(dev) if k != 'planes' else v) for k, v in inp.items()}
pl = rmod.planes_channel_minor(dd['planes']).requires_grad_(True)
ws = [dd[k].clone().requires_grad_(True) for k in ('w1', 'b1', 'w2', 'b2')]
c2w = ru.compute_cam2world_matrix(dn.TensorGroup(angles=dd['angles'], radius=torch.ones(B, device=dev), look_at=dd['look_at']))
rgb, depth, _, _ = rmod.render_camera(pl, *ws, c2w, dd['fov'], (64, 64), num_steps=48, ray_start=0.75, ray_end=1.25, box_size=1.0, density_noise=0.5, seed=1, mlp_mode=2)
torch.autograd.grad([rgb, depth], [pl] + ws, [torch.ones_like(rgb), torch.ones_like(depth)])
torch.cuda.synchronize()
print('done')
