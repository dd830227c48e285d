"""Kernel-time breakdown of one training step (torch.profiler / CUPTI).  python tools/profile_step.py [batch] [out]"""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
out = sys.argv[2] if len(sys.argv) > 2 else 'gpurun_out/step_profile.txt'
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
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step(real, gen)      # it=3: no Dreg
    torch.cuda.synchronize()
tab = prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=90)
open(out, 'w').write(tab)
print(tab[:6000])
