"""Where does the ray-march forward's time go?  Same launch geometry (B=16, 64x64 rays, 48+48 samples), one factor changed at a time:
MLP arithmetic (3xTF32 / TF32), plane resolution (512^2: L2 gather; 64^2: the whole plane set sits in L1/L2 -- gather latency removed), plane storage
(fp32 / fp16), density noise (Philox on / off).  python tools/rm_probe.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
rm = importlib.import_module('3dgp_b200.torch_utils.ops.raymarch')
dn = importlib.import_module('3dgp_b200.dnnlib'); ru = importlib.import_module('3dgp_b200.training.rendering_utils')
dev = torch.device('cuda')
B = 16


def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n


inp = bench.raymarch_inputs(B, dev, 0)
d = {k: (v.to(dev) if k != 'planes' else v) for k, v in inp.items()}
c2w = ru.compute_cam2world_matrix(dn.TensorGroup(angles=d['angles'], radius=torch.ones(B, device=dev), look_at=d['look_at']))
for P in (512, 64):
    planes = torch.randn([B, P, P, 96], device=dev).permute(0, 3, 1, 2).view(B, 3, 32, P, P)
    pl = rm.planes_channel_minor(planes)
    for half in (False, True):
        plx = pl.half() if half else pl
        for mode in (2, 1):
            for noise in (0.0, 0.5):
                kw = dict(num_steps=48, ray_start=0.75, ray_end=1.25, box_size=1.0, mlp_mode=mode, density_noise=noise, seed=1)
                ms = t(lambda: rm.render_camera(plx, d['w1'], d['b1'], d['w2'], d['b2'], c2w, d['fov'], (64, 64), **kw))
                print(f'P={P:3d} planes={"f16" if half else "f32"} mlp_mode={mode} density_noise={noise}: {ms:.3f} ms', flush=True)
