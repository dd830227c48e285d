#!/bin/bash
# Round-2 GPU pass B: third-generation ray-march forward (parity, bench, ncu), wide-network parity, conv traffic over whole steps.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py -m gpu -x -q > gpurun_out/b_pytest_render.log 2>&1; echo "rc=$?" >> gpurun_out/b_pytest_render.log
timeout 900 python -m pytest tests/test_gpu_networks_wide.py tests/test_gpu_networks.py -m gpu -q > gpurun_out/b_pytest_nets.log 2>&1; echo "rc=$?" >> gpurun_out/b_pytest_nets.log
timeout 300 python bench.py --workload raymarch --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_rm.json 2> gpurun_out/b_bench_rm.err
GP3D_RAYMARCH_V2=1 timeout 300 python - > gpurun_out/b_rm_v2_vs_v3.txt 2>&1 <<'PY'
import importlib, sys, torch, numpy as np
sys.path.insert(0, '.')
import bench
rm = importlib.import_module('3dgp_b200.torch_utils.ops.raymarch')
dn = importlib.import_module('3dgp_b200.dnnlib'); ru = importlib.import_module('3dgp_b200.training.rendering_utils')
dev = torch.device('cuda')
B = 16
inp = bench.raymarch_inputs(B, dev, 0)
d = {k: (v.to(dev) if k != 'planes' else v) for k, v in inp.items()}
pl = rm.planes_channel_minor(d['planes'])
c2w = ru.compute_cam2world_matrix(dn.TensorGroup(angles=d['angles'], radius=torch.ones(B, device=dev), look_at=d['look_at']))
ro, rd = rm.generate_rays(c2w, d['fov'], (64, 64))
kw = dict(num_steps=48, ray_start=0.75, ray_end=1.25, box_size=1.0, mlp_mode=2)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
print('v2 explicit rays (env GP3D_RAYMARCH_V2=1) ms', t(lambda: rm.render_rays(pl, d['w1'], d['b1'], d['w2'], d['b2'], ro, rd, seed=1, **kw)))
print('v3 camera tiles 4x4 ms', t(lambda: rm.render_camera(pl, d['w1'], d['b1'], d['w2'], d['b2'], c2w, d['fov'], (64, 64), seed=1, **kw)))
a = rm.render_rays(pl, d['w1'], d['b1'], d['w2'], d['b2'], ro, rd, seed=1, **kw)
b = rm.render_camera(pl, d['w1'], d['b1'], d['w2'], d['b2'], c2w, d['fov'], (64, 64), seed=1, **kw)
print('max abs diff v2 vs v3', [float((x - y).abs().max()) for x, y in zip(a, b)])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raymarch_fwd3 -s 3 -c 1 -o gpurun_out/b_rm_fwd3 python bench.py --workload raymarch --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_rm.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_nhwc_bf16_kernel -c 1600 --csv --log-file gpurun_out/b_conv_traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-gpu-baseline > gpurun_out/b_ncu_bench.log 2>&1
echo done
