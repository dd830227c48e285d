"""Ray-march forward / forward+backward timing at BASELINE config 3 (B=16, 64x64 rays, 48+48 samples).  python tools/rm_bwd_probe.py [modes] [reps]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
modes = [int(m) for m in (sys.argv[1] if len(sys.argv) > 1 else '0,1,2').split(',')]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rmod = importlib.import_module('3dgp_b200.torch_utils.ops.raymarch')
dev = torch.device('cuda')
B = 16
inp = bench.raymarch_inputs(B, dev, seed=0)
d = {k: (v.to(dev) if k != 'planes' else v) for k, v in inp.items()}
pl = rmod.planes_channel_minor(d['planes']).requires_grad_(True)
ws = [d[k].clone().requires_grad_(True) for k in ('w1', 'b1', 'w2', 'b2')]
kw = dict(num_steps=48, ray_start=0.75, ray_end=1.25, box_size=1.0, density_noise=0.5)
for mode in modes:
    def fb(seed):
        rgb, depth, _, _ = rmod.render_rays(pl, *ws, d['ray_o'], d['ray_d'], seed=seed, mlp_mode=mode, **kw)
        g = torch.autograd.grad([rgb, depth], [pl] + ws, [torch.ones_like(rgb), torch.ones_like(depth)])
        return g
    for i in range(2):
        fb(i)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for i in range(reps):
        e[0].record()
        rgb, depth, _, _ = rmod.render_rays(pl, *ws, d['ray_o'], d['ray_d'], seed=i, mlp_mode=mode, **kw)
        e[1].record()
        g = torch.autograd.grad([rgb, depth], [pl] + ws, [torch.ones_like(rgb), torch.ones_like(depth)])
        e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    print(f'mode {mode}: forward {tf / reps:.3f} ms   backward (incl. 1.6 GB zero-fill of g_planes) {tb / reps:.3f} ms', flush=True)
