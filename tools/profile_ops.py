"""Operator-level (with input shapes) breakdown of one training step: which torch ops still run outside lib3dgp_b200."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
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
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    tr.step(real, gen)
    torch.cuda.synchronize()
ka = prof.key_averages(group_by_input_shape=True)
rows = []
for e in ka:
    t = getattr(e, 'self_device_time_total', None)
    if t is None:
        t = getattr(e, 'self_cuda_time_total', 0)
    if t > 300 and e.key.startswith('aten::'):
        rows.append((t / 1e3, e.count, e.key, str(e.input_shapes)[:150]))
rows.sort(reverse=True)
out = '\n'.join(f'{t:9.3f} ms  x{c:<4d} {k:28s} {s}' for t, c, k, s in rows[:60])
open('gpurun_out/ops_by_shape.txt', 'w').write(out)
print(out)
