"""Kernel-time breakdown of the lazy-regularisation iteration (Gmain + Dmain + Dreg).  python tools/profile_dreg.py [batch] [out]"""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
out = sys.argv[2] if len(sys.argv) > 2 else 'gpurun_out/dreg_profile.txt'
cfgm = importlib.import_module('3dgp_b200.config'); dn = importlib.import_module('3dgp_b200.dnnlib')
lossm = importlib.import_module('3dgp_b200.training.loss'); stepm = importlib.import_module('3dgp_b200.training.step')
dev = torch.device('cuda')
cfg = cfgm.make_config(batch_size=B)
torch.manual_seed(0); np.random.seed(0)
G, D = cfgm.build_networks(cfg, dev)
G.train(); D.train()
loss = lossm.StyleGAN2Loss(cfg, dev, G, D, r1_gamma=0.8)
tr = stepm.Trainer(G, D, loss, cfg, D_reg_interval=16, batch_size=B, micro_batch=B)
host = bench.synthetic_batch(cfg, B, dev, 0)
real, gen = bench.to_step_inputs(host, dev, dn)
for _ in range(3):
    tr.step(real, gen)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record(); tr._phase('Dreg', tr.D, tr.D_opt, real, gen, 16); ev[1].record(); torch.cuda.synchronize()
print('Dreg phase alone: %.1f ms' % ev[0].elapsed_time(ev[1]))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    tr._phase('Dreg', tr.D, tr.D_opt, real, gen, 16)
    torch.cuda.synchronize()
tab = prof.key_averages().table(sort_by='cuda_time_total', row_limit=40, max_name_column_width=90)
open(out, 'w').write(tab)
evs = {}
for e in prof.key_averages(group_by_input_shape=True):
    if e.device_time_total > 0 and e.key.startswith('aten::'):
        evs[(e.key, str(e.input_shapes)[:90])] = (e.device_time_total / 1e3, e.count)
print(tab[:100])
for (k, sh), (t, c) in sorted(evs.items(), key=lambda kv: -kv[1][0])[:25]:
    print('%9.3f ms  x%-4d %-34s %s' % (t, c, k, sh))
