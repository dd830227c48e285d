#!/usr/bin/env python
"""bench.py -- one JSON line per run (driver contract, see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train_step|raymarch|ginfer|ops] [--impl ours|reference]

Workloads (BASELINE.json `configs`):
  train_step : configs[1]  ImageNet-256 G+D training step (Gmain + Dmain phases), cmax=1024, 48 samples/ray,
               patch 64x64, batch/GPU from --batch-gpu, synthetic data, random-init weights.  metric = images/s.
  raymarch   : configs[2]  fused ray-march microbench, 64x64 rays, 48+48 samples/ray, 32-ch 512^2 tri-planes, batch 16.
  ginfer     : configs[4]  G-only inference at 256x256.
`--impl reference` EXECUTES the reference's CPU algorithm (the oracle port, torch-CPU, all host threads: forward, backward and optimiser
of every phase) for W + K steps on a bounded sample of the same workload; under torchrun only rank 0 runs it.

Timing: W >= 3 warm-up steps, then exactly K steps between barrier + cuda synchronize, CUDA events on the launching
(torch current) stream, max over ranks.  Inputs are larger than L2 (stated in config.l2).  Clocks are sampled with
nvidia-smi during the timed region.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d['hbm_gbs']), bf16_tflops=float(d['bf16_tflops']),
                    bf16_tflops_sustained=float(d.get('bf16_tflops_sustained', d['bf16_tflops'])), source='measured')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source='fallback')   # B200_PROFILING.md


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.index = index; self.rows = []; self.proc = None; self.th = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100', '-i', str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def rd():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.th = threading.Thread(target=rd, daemon=True); self.th.start()

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons), samples=len(sm))


def dist_info():
    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1)); local = int(os.environ.get('LOCAL_RANK', 0))
    return rank, world, local


def timed_region(step_fn, steps, warmup, world):
    """Returns (ms_per_step max over ranks, clocks dict).  step_fn() enqueues one step on the current stream."""
    import torch.distributed as dist
    for _ in range(max(warmup, 3)):
        step_fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(torch.cuda.current_device()); sampler.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    return ms / steps, clocks


# ----------------------------------------------------------------------------------------------
# workload: fused ray-march microbench (BASELINE configs[2])
RM = dict(B=16, P=512, C=32, H=64, res=64, N=48, ray_start=0.75, ray_end=1.25, box_half=0.5)


def raymarch_inputs(B, device, seed=0):
    """Synthetic tri-planes N(0,1) stored channel-minor (the layout the tri-plane decoder emits), cameras ~ configs/camera/{base,uniform}.yaml
    (full-frame 64x64 rays), random-init MLP (layers.py:36: randn weights, zero bias).  Returns camera PARAMETERS; rays are generated by the consumer:
    inside the kernel on the GPU arm, by the oracle's restated ray generator on the CPU arm."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    P, C = RM['P'], RM['C']
    if torch.device(device).type == 'cuda':
        planes = torch.randn([B, P, P, 3 * C], device=device, generator=torch.Generator(device=device).manual_seed(seed))
    else:
        planes = torch.randn([B, P, P, 3 * C], generator=torch.Generator().manual_seed(seed))
    planes = planes.permute(0, 3, 1, 2).view(B, 3, C, P, P)
    yaw = torch.rand(B, generator=g) * 3.14 - 1.57
    pitch = torch.rand(B, generator=g) * (2.35619449 - 0.785398163) + 0.785398163
    angles = torch.stack([yaw, pitch, torch.zeros(B)], 1)
    fov = torch.rand(B, generator=g) * 35 + 10
    look = torch.stack([torch.rand(B, generator=g) * 6.28 - 3.14, torch.acos(1 - 2 * torch.rand(B, generator=g).clamp(1e-5, 1 - 1e-5)), torch.rand(B, generator=g) * 0.2], 1)
    w1 = torch.randn(RM['H'], C, generator=g); w2 = torch.randn(4, RM['H'], generator=g)
    return dict(planes=planes, angles=angles, fov=fov, look_at=look, w1=w1, b1=torch.zeros(RM['H']), w2=w2, b2=torch.zeros(4))


def raymarch_algorithmic_bytes(B, R, N, P, C, plane_bytes=4, injected_noise=False):
    """planes once + rays in (+ injected variates) + results out."""
    return B * 3 * C * P * P * plane_bytes + B * R * (24 + (2 * N * 4 if injected_noise else 0)) + B * R * 24


def run_raymarch(args, rank, world, local):
    rmod = importlib.import_module('3dgp_b200.torch_utils.ops.raymarch')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    ru = importlib.import_module('3dgp_b200.training.rendering_utils')
    dev = torch.device('cuda', local)
    B = args.batch_gpu or RM['B']
    inp = raymarch_inputs(B, dev, seed=rank)
    R_ = RM['res'] ** 2
    res = (RM['res'], RM['res'])
    d = {k: (v.to(dev) if k != 'planes' else v) for k, v in inp.items()}
    kw = dict(num_steps=RM['N'], ray_start=RM['ray_start'], ray_end=RM['ray_end'], box_size=2 * RM['box_half'], mlp_mode=args.mlp_mode)
    launches = {'n': 0}
    pl = rmod.planes_channel_minor(d['planes'])
    if args.planes_fp16:
        pl = pl.half()
    c2w = ru.compute_cam2world_matrix(dn.TensorGroup(angles=d['angles'], radius=torch.ones(B, device=dev), look_at=d['look_at']))

    def step():
        out = rmod.render_camera(pl, d['w1'], d['b1'], d['w2'], d['b2'], c2w, d['fov'], res, seed=launches['n'], **kw)
        launches['n'] += 1
        return out

    rmod.TIMING = None
    ms, clocks = timed_region(step, args.steps, args.warmup, world)
    # the kernel alone, live: CUDA events around every launch on the launching stream
    rmod.TIMING = []
    timed_region(step, args.steps, 0, world)
    ev = rmod.TIMING; rmod.TIMING = None
    torch.cuda.synchronize()
    kms = float(np.mean([a.elapsed_time(b) for (a, b, *_r) in ev]))
    # end-to-end: camera parameters from pinned host memory every step (cam2world is built on the device), results read back to the host
    cam_h = {k: inp[k].pin_memory() for k in ('angles', 'fov', 'look_at')}
    out_h = torch.empty([B, R_, 5], pin_memory=True)

    def step_e2e():
        cd = {k: v.to(dev, non_blocking=True) for k, v in cam_h.items()}
        c2 = ru.compute_cam2world_matrix(dn.TensorGroup(angles=cd['angles'], radius=torch.ones(B, device=dev), look_at=cd['look_at']))
        rgb, depth, wsum, _ = rmod.render_camera(pl, d['w1'], d['b1'], d['w2'], d['b2'], c2, cd['fov'], res, seed=launches['n'], **kw)
        launches['n'] += 1
        out_h.copy_(torch.cat([rgb, depth, wsum], dim=-1), non_blocking=True)

    ms_e2e, _ = timed_region(step_e2e, args.steps, args.warmup, world)
    # config 3 also asks for forward + backward: gradients w.r.t. planes and MLP (the training configuration), CUDA events per leg
    plg = pl.detach().clone().requires_grad_(True)
    wsg = [d[k].clone().requires_grad_(True) for k in ('w1', 'b1', 'w2', 'b2')]
    tf = tb = 0.0
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for i in range(args.warmup + args.steps):
        evs[0].record()
        rgb, depth, _, _ = rmod.render_camera(plg, *wsg, c2w, d['fov'], res, seed=i, density_noise=0.5, **kw)
        evs[1].record()
        torch.autograd.grad([rgb, depth], [plg] + wsg, [torch.ones_like(rgb), torch.ones_like(depth)])
        evs[2].record()
        torch.cuda.synchronize()
        if i >= args.warmup:
            tf += evs[0].elapsed_time(evs[1]); tb += evs[1].elapsed_time(evs[2])
    del plg
    pb = 2 if args.planes_fp16 else 4
    alg = raymarch_algorithmic_bytes(B, R_, RM['N'], RM['P'], RM['C'], plane_bytes=pb)
    peaks = measured_peaks()
    ach = alg / (kms * 1e-3) / 1e9
    traffic = RAYMARCH_TRAFFIC.get(f'B{B}_mode{args.mlp_mode}_{"f16" if args.planes_fp16 else "f32"}')
    res_ = dict(
        metric='ray-march images/s (64x64 rays, 48+48 samples/ray, 32-ch 512^2 tri-planes)', value=world * B / (ms * 1e-3), unit='images/s',
        ms_per_step=ms, dtype='f32 (3xTF32 mma.sync MLP)' if not args.planes_fp16 else 'f32 math / f16 planes',
        config=dict(workload='raymarch (BASELINE configs[2])', batch_per_gpu=B, rays=R_, samples_per_ray=2 * RM['N'], plane_res=RM['P'],
                    l2='inputs (%.2f GB of planes per GPU) larger than the 126 MB L2' % (B * 3 * RM['C'] * RM['P'] ** 2 * pb / 1e9),
                    rng='in-kernel Philox', rays_from='camera, generated in the kernel', mlp_mode=args.mlp_mode, parallelism=f'replicas x{world} (render does not shard)'),
        roofline=dict(bound='hbm', achieved=ach, peak=peaks['hbm_gbs'], unit='GB/s', frac=ach / peaks['hbm_gbs'], traffic=traffic,
                      kernel='raymarch_fwd3_kernel' if args.mlp_mode else 'raymarch_fwd_kernel', peak_source=peaks['source'] + ' (burst copy: kernel timed alone)',
                      algorithmic_bytes_per_launch=alg, mean_kernel_ms=kms, launches_timed=len(ev),
                      traffic_source='profiles/raymarch_traffic.json (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum of one launch)' if traffic else None),
        e2e=dict(value=world * B / (ms_e2e * 1e-3), unit='images/s', h2d_bytes_per_step=int(sum(v.numel() * 4 for v in cam_h.values())), d2h_bytes_per_step=int(B * R_ * 20)),
        gpu_launches=args.steps, clocks=clocks,
        forward_backward=dict(forward_ms=tf / args.steps, backward_ms=tb / args.steps, kernel_bwd='raymarch_bwd2_kernel' if args.mlp_mode else 'raymarch_bwd_kernel',
                              note='backward includes the ray generator launch and the zero-fill of the plane-gradient buffer (B*3*C*P^2*4 bytes)'))
    return res_


RAYMARCH_TRAFFIC = {}
try:
    RAYMARCH_TRAFFIC = json.load(open(os.path.join(ROOT, 'profiles', 'raymarch_traffic.json')))
except Exception:
    pass


def cpu_raymarch(sample_rays=4096, repeats=2):
    """Reference algorithm on the host cores (oracle port), bounded sample: one image, `sample_rays` rays."""
    from oracle import restated as R
    torch.set_num_threads(os.cpu_count())
    inp = raymarch_inputs(1, 'cpu', seed=0)
    planes = inp['planes'].contiguous()
    c2w = R.compute_cam2world_matrix(inp['angles'], torch.ones(1), inp['look_at'])
    ro, rd = R.sample_rays(c2w, inp['fov'], (RM['res'], RM['res']))
    ro = ro[:, :sample_rays]; rd = rd[:, :sample_rays]
    N = RM['N']
    u1 = torch.rand(1, sample_rays, N); u2 = torch.rand(1, sample_rays, N)
    best = None
    for _ in range(repeats + 1):
        t0 = time.perf_counter()
        R.render(planes, inp['w1'], inp['b1'], inp['w2'], inp['b2'], ro, rd, u1, u2, RM['ray_start'], RM['ray_end'], RM['box_half'], N)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    imgs_per_s = (sample_rays / RM['res'] ** 2) / best
    return dict(value=imgs_per_s, unit='images/s', cores=os.cpu_count(), kind='port',
                sample=f'1 image x {sample_rays} of 4096 rays x 96 samples, best of {repeats + 1}, {best:.2f} s')


# ----------------------------------------------------------------------------------------------
# workload: G+D training step (BASELINE configs[1])
def synthetic_batch(cfg, B, device, seed):
    """Synthetic data of the ImageNet-256 shape: uint8-like images, 16-bit-like depths, one-hot labels,
    ResNet50-like 2048-d embeddings, latents, cameras from the uniform prior.  Returned on the HOST (pinned)."""
    g = torch.Generator().manual_seed(seed)
    res = cfg.dataset.resolution
    img = torch.randint(0, 256, (B, 3, res, res), generator=g).float() / 127.5 - 1
    depth = torch.randint(0, 65536, (B, 1, res, res), generator=g).float() / 65536 * 2 - 1
    c = torch.zeros(B, cfg.dataset.c_dim); c[torch.arange(B), torch.arange(B) % cfg.dataset.c_dim] = 1
    embs = torch.randn(B, cfg.dataset.embedding_dim, generator=g)
    z = torch.randn(B, cfg.model.generator.z_dim, generator=g)
    yaw = torch.rand(B, generator=g) * 3.14 - 1.57
    pitch = torch.rand(B, generator=g) * (2.35619449 - 0.785398163) + 0.785398163
    angles = torch.stack([yaw, pitch, torch.zeros(B)], 1)
    fov = torch.rand(B, generator=g) * 35 + 10
    look = torch.stack([torch.rand(B, generator=g) * 6.28 - 3.14, torch.acos(1 - 2 * torch.rand(B, generator=g).clamp(1e-5, 1 - 1e-5)), torch.rand(B, generator=g) * 0.2], 1)
    host = dict(img=img, depth=depth, c=c, embs=embs, z=z, angles=angles, fov=fov, radius=torch.ones(B), look_at=look)
    return {k: v.pin_memory() for k, v in host.items()}


def to_step_inputs(host, device, dn):
    d = {k: v.to(device, non_blocking=True) for k, v in host.items()}
    real = dn.EasyDict(img=d['img'], depth=d['depth'], c=d['c'], embs=d['embs'], camera_angles=d['angles'])
    gen = dn.EasyDict(z=d['z'], c=d['c'], camera_params=dn.TensorGroup(angles=d['angles'], fov=d['fov'], radius=d['radius'], look_at=d['look_at']))
    return real, gen


def conv_flops_per_image(cfg):
    """Dense-contraction FLOPs of one image through G (decoder) and D, forward (492.0 / 209.8 GFLOP at the BASELINE config)."""
    g = cfg.model.generator
    res_list = [2 ** i for i in range(2, int(np.log2(g.tri_plane.res)) + 1)]
    ch = {r: min(int(g.cbase * g.fmaps) // r, g.cmax) for r in res_list}
    oc = 3 * g.tri_plane.feat_dim
    fl = 0
    for r in res_list:
        if r > 4:
            fl += 2 * ch[r // 2] * ch[r] * 9 * (r // 2) ** 2       # conv0, transpose-conv formulation (what runs)
        fl += 2 * ch[r] * ch[r] * 9 * r * r                         # conv1
        fl += 2 * ch[r] * oc * r * r                                # torgb
    d = cfg.model.discriminator
    pr = cfg.training.patch.resolution
    top = pr * 2 ** d.num_additional_start_blocks
    dres = [2 ** i for i in range(int(np.log2(top)), 2, -1)]
    dch = {r: min(int(d.cbase * d.fmaps) // r, d.cmax) for r in dres + [4]}
    fd = 0
    hw = pr
    for i, r in enumerate(dres):
        down = 1 if i < d.num_additional_start_blocks else 2
        if i == 0:
            fd += 2 * 4 * dch[r] * hw * hw
        fd += 2 * dch[r] * dch[r] * 9 * hw * hw                     # conv0
        out_hw = hw // down
        fd += 2 * dch[r] * dch[r // 2] * 9 * out_hw * out_hw        # conv1 (stride `down`)
        fd += 2 * dch[r] * dch[r // 2] * out_hw * out_hw            # skip 1x1
        hw = out_hw
    fd += 2 * (dch[4] + 1) * dch[4] * 9 * 16 + 2 * dch[4] * 16 * dch[4] * 2 + 2 * dch[4] * dch[4]
    return fl, fd


# ncu --set full capture of the step's dominant kernel (profiles/, see DESIGN.md "Measurement"): dram__bytes_read.sum + dram__bytes_write.sum per launch,
# averaged over the launches of conv_nhwc_bf16_kernel<128,3> inside one training step at the benchmarked configuration; None until measured for a config.
CONV_TRAFFIC_BYTES_PER_LAUNCH = {}
try:
    CONV_TRAFFIC_BYTES_PER_LAUNCH = json.load(open(os.path.join(ROOT, 'profiles', 'conv_traffic_per_launch.json')))
except Exception:
    pass

TRAIN_WORKLOAD = 'train_step (BASELINE configs[1]: ImageNet-256 G+D step, cmax=1024, 48+48 samples/ray, patch 64x64, learn_camera_dist=true)'


def train_step_config(B, mb, world, small=False):
    """The `config` object of the JSON line: static description of the workload, identical for the GPU arm and the `--impl reference` arm."""
    return dict(workload=TRAIN_WORKLOAD if not small else 'train_step SMALL (debug)', batch_per_gpu=B, micro_batch=mb, global_batch=B * world,
                phases='Gmain + Dmain every iteration, Dreg (R1) every 16th', learn_camera_dist=True,
                l2='activations per layer (>= 134 MB/image at 512^2) larger than the 126 MB L2',
                parallelism=f'dp{world}: one flattened gradient all-reduce per phase (NCCL)')


def run_train_step(args, rank, world, local):
    gp = importlib.import_module('3dgp_b200')
    cfgm = importlib.import_module('3dgp_b200.config')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    lossm = importlib.import_module('3dgp_b200.training.loss')
    stepm = importlib.import_module('3dgp_b200.training.step')
    rmod = importlib.import_module('3dgp_b200.torch_utils.ops.raymarch')
    tcm = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
    dev = torch.device('cuda', local)
    B = args.batch_gpu or 32
    mb = args.micro_batch or min(B, 32)
    assert B % mb == 0 and mb % 4 == 0, 'micro-batch must divide batch-gpu and be a multiple of the minibatch-std group (4)'
    small = dict(cmax=64, cbase=4096, tri_res=128, patch_res=32, img_resolution=128, c_dim=10, w_dim=128, z_dim=128, num_ray_steps=12) if args.small else {}
    cfg = cfgm.make_config(batch_size=B * world, learn_camera_dist=True, **small)        # configs/training/base.yaml:9 (the reference default)
    torch.manual_seed(1234 + rank); np.random.seed(1234 + rank)
    G, D = cfgm.build_networks(cfg, dev)
    with torch.no_grad():   # benchmark init: exercise the noise path
        for n, p_ in G.named_parameters():
            if n.endswith('noise_strength'):
                p_.fill_(0.1)
    G.train(); D.train()
    r1_gamma = 0.0002 * (cfg.dataset.resolution ** 2) / (B * world)      # train.py:173 'auto'
    loss = lossm.StyleGAN2Loss(cfg, dev, G, D, r1_gamma=r1_gamma)
    # mid-training state (2500 kimg): density noise at half strength, EMD regulariser of the camera adaptor ramped in (loss.py:64-65)
    G.progressive_update(2500); loss.progressive_update(2500)
    tr = stepm.Trainer(G, D, loss, cfg, rank=rank, world_size=world, D_reg_interval=16, batch_size=B * world, micro_batch=mb)
    tr.overlap_allreduce = bool(args.overlap)
    host = synthetic_batch(cfg, B, dev, seed=rank)
    real, gen = to_step_inputs(host, dev, dn)
    torch.cuda.synchronize()

    def step():
        tr.step(real, gen)

    rmod.TIMING = None
    ms, clocks = timed_region(step, args.steps, args.warmup, world)
    # same steps, collecting the durations of the dominant conv kernel and of the fused ray-march forward with events on the launching stream
    c0 = gp._lib.launch_count
    rmod.TIMING = []
    tcm.CONV_TIMING = []
    s0 = dict(tcm.stats)
    ms2, _ = timed_region(step, args.steps, 0, world)
    launches = (gp._lib.launch_count - c0) // max(args.steps + 3, 1)
    conv_routing = {k: (tcm.stats[k] - s0[k]) // max(args.steps + 3, 1) for k in s0}
    ev = rmod.TIMING; rmod.TIMING = None
    cev = tcm.CONV_TIMING; tcm.CONV_TIMING = None
    torch.cuda.synchronize()
    conv_ms = float(sum(a.elapsed_time(b) for a, b, _ in cev)) or 1e-9
    conv_flops = float(sum(f_ for _, _, f_ in cev))
    kms = [a.elapsed_time(b) for (a, b, *_rest) in ev]
    _, _, Bk, Rk, Nk, Pk, Ck, esz = ev[0]
    alg = raymarch_algorithmic_bytes(Bk, Rk, Nk, Pk, Ck, plane_bytes=esz)
    peaks = measured_peaks()
    ach = alg / (float(np.mean(kms)) * 1e-3) / 1e9

    # end-to-end through the public API: every step copies its inputs from pinned host memory and reads the losses back
    stats_h = torch.empty(4, pin_memory=True)

    def step_e2e():
        r, g_ = to_step_inputs(host, dev, dn)
        st = tr.step(r, g_)
        vals = torch.stack([st.get('Loss/G/loss', torch.zeros((), device=dev)), st.get('Loss/scores/fake', torch.zeros((), device=dev)),
                            st.get('Loss/scores/real', torch.zeros((), device=dev)), st.get('Loss/D/r1_penalty', torch.zeros((), device=dev))])
        stats_h.copy_(vals, non_blocking=True)

    ms_e2e, _ = timed_region(step_e2e, args.steps, 1, world)
    peak_mem = round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    fg, fd = conv_flops_per_image(cfg)
    # per iteration: Gmain = G fwd+bwd (3x) + D fwd + dgrad (2x); Dmain = G fwd (1x) + D fwd+bwd on fakes and reals (2 x 3x); Dreg/16 ~ D 2nd order
    flops_step = B * (4 * fg + 8 * fd + (1.0 / 16) * 6 * fd)
    alg_tf = conv_flops / (conv_ms * 1e-3) / 1e12
    n_l = max(len(cev), 1)
    traffic = CONV_TRAFFIC_BYTES_PER_LAUNCH.get(f'B{mb}')
    # The comparison legs below are reported next to the headline, never instead of it: a failure inside one of them (say, cuDNN running out of workspace
    # on the unfused path) is recorded under its key and the measured line is still printed.
    gpu_base = None
    if world == 1 and not args.no_gpu_baseline and not args.small:
        gpu_base = _side_leg('gpu_baseline', gpu_baseline_step, tr, host, dev, dn, args)
    ginfer = None
    if world == 1 and not args.no_ginfer and not args.small:      # replicas only: measured at N = 1 (use --workload ginfer under torchrun for N > 1)
        ginfer = _side_leg('ginfer', ginfer_leg, tr.G_ema, cfg, dev, dn, rank, world)
    res = dict(
        metric='G+D training-step images/s at 256x256', value=world * B / (ms * 1e-3), unit='images/s', ms_per_step=ms,
        dtype='f32 storage; G convs bf16x3 on tcgen05 (three bf16 MMAs per product, fp32 accumulate: fp32-grade), tri-plane MLP 3xTF32; '
              'D blocks the reference runs in fp16: %s tensor-core operands, fp32 accumulate' % D_LOW_PRECISION_NAME,
        config=train_step_config(B, mb, world, args.small),
        details=dict(peak_mem_gb=peak_mem, conv_gflop_per_image_fwd=dict(G=fg / 1e9, D=fd / 1e9), conv_routing_per_step=conv_routing,
                     density_noise_std=float(G.synthesis.nerf_noise_std), emd_multiplier=float(loss.emd_multiplier)),
        # dominant kernel of the step: the stride-1 bf16x3 tcgen05 convolution of the tri-plane decoder (forward + input gradient).
        # `achieved` = ALGORITHMIC convolution FLOPs (2*B*Cout*Cin*k^2*H*W) / live-timed launch duration.
        roofline=dict(bound='tensor', achieved=alg_tf, peak=peaks['bf16_tflops_sustained'], unit='TFLOP/s', frac=alg_tf / peaks['bf16_tflops_sustained'],
                      traffic=traffic, kernel='conv_nhwc_bf16_kernel<*,3> (three-term bf16x3 form: 256-wide tiles where Cout % 256 == 0, else 128 / 96)', peak_source=peaks['source'] + ' (sustained bf16 GEMM: kernel timed inside a long step)',
                      launches_timed=len(cev), mean_ms=conv_ms / n_l, algorithmic_flops_per_launch=conv_flops / n_l,
                      # bf16x3 executes three bf16 MMAs per fp32-grade product: tensor-pipe occupancy, NOT the roofline fraction
                      executed_mma_tflops=3.0 * alg_tf, executed_mma_frac_of_peak=3.0 * alg_tf / peaks['bf16_tflops_sustained'],
                      traffic_source='profiles/conv_traffic_per_launch.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean per launch)' if traffic else None),
        roofline_raymarch=dict(bound='hbm', achieved=ach, peak=peaks['hbm_gbs'], unit='GB/s', frac=ach / peaks['hbm_gbs'], traffic=None,
                               kernel='raymarch_fwd kernel (3xTF32 MLP, inside the step)', peak_source=peaks['source'], algorithmic_bytes_per_launch=alg,
                               launches_timed=len(kms), mean_ms=float(np.mean(kms))),
        roofline_step_tensor=dict(bound='tensor', achieved=flops_step / (ms * 1e-3) / 1e12, peak=peaks['bf16_tflops_sustained'], unit='TFLOP/s',
                                  frac=flops_step / (ms * 1e-3) / 1e12 / peaks['bf16_tflops_sustained'],
                                  note='algorithmic dense-contraction FLOPs of the whole step / step time, against the measured sustained bf16 GEMM peak'),
        e2e=dict(value=world * B / (ms_e2e * 1e-3), unit='images/s', h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=16),
        gpu_launches=int(launches * args.steps), clocks=clocks)
    if gpu_base is not None:
        res['gpu_baseline'] = gpu_base
    if ginfer is not None:
        res['ginfer'] = ginfer
    return res


def _side_leg(name, fn, *a, **k):
    """Runs a comparison leg; on failure returns {'error': ...} (traceback on stderr) instead of taking the headline measurement down with it."""
    try:
        return fn(*a, **k)
    except Exception as e:          # noqa: BLE001 -- any failure of a side leg is reported, not raised
        import traceback
        traceback.print_exc(file=sys.stderr)
        print(f'bench.py: side leg {name} failed: {e!r}', file=sys.stderr)
        try:
            import torch
            if torch.cuda.is_available():
                torch.cuda.empty_cache()
        except Exception:
            pass
        return dict(error=f'{type(e).__name__}: {e}'[:300])


D_LOW_PRECISION_NAME = 'fp16 x fp16 forward products (bf16 x bf16 for products with a gradient operand)'


def ginfer_leg(G_ema, cfg, dev, dn, rank, world, B=64, steps=3):
    """BASELINE configs[4] inside the default run: G_ema inference at batch 64 through the metrics loop's call pattern (metric_utils.py:303-319,
    training/inference.py::generate_uint8: eval mode, 256x256 full-frame render, uint8 on the device), every step copying its latents / cameras from
    pinned host memory and the uint8 images back.  Replicas only: no collective."""
    inf = importlib.import_module('3dgp_b200.training.inference')
    host = synthetic_batch(cfg, B, dev, seed=1000 + rank)
    keys = ('z', 'c', 'angles', 'fov', 'radius', 'look_at')
    res = cfg.dataset.resolution
    out_h = torch.empty([B, 3, res, res], dtype=torch.uint8, pin_memory=True)

    def step():
        dd = {k: host[k].to(dev, non_blocking=True) for k in keys}
        cm = dn.TensorGroup(angles=dd['angles'], fov=dd['fov'], radius=dd['radius'], look_at=dd['look_at'])
        out_h.copy_(inf.generate_uint8(G_ema, dd['z'], dd['c'], cm, noise_mode='const'), non_blocking=True)

    ms, _ = timed_region(step, steps, 3, world)
    return dict(metric='G-only inference images/s at 256x256 (end to end: H2D latents / cameras, D2H uint8 images)', value=world * B / (ms * 1e-3), unit='images/s',
                ms_per_step=ms, batch_per_gpu=B, steps=steps, rays_per_image=res * res, d2h_bytes_per_step=int(out_h.numel()),
                parallelism=f'replicas x{world} (no collective)')


def gpu_baseline_step(tr, host, dev, dn, args):
    """Same modules, same weights, same step -- but every convolution on cuDNN fp32 with TF32 off (training_loop.py:76-77, the reference's own GPU
    arithmetic for G) and every fused layer node switched off, i.e. the op-by-op composition the reference executes (x*styles -> conv -> FIR -> fma ->
    bias_act; hyper-mod -> conv -> bias_act), on this repo's elementwise / FIR / ray-march kernels.  Timed in the same run on a reduced batch (the unfused
    path keeps every intermediate alive for backward: 32 images do not fit comfortably)."""
    cg = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix')
    sg = importlib.import_module('3dgp_b200.training.networks_stylegan2')
    layers = importlib.import_module('3dgp_b200.training.layers')
    Bb = 8
    sub = {k: v[:Bb] for k, v in host.items()}
    real, gen = to_step_inputs(sub, dev, dn)
    old_mb, old_bs = tr.micro_batch, tr.batch_size
    tr.micro_batch, tr.batch_size = Bb, Bb
    cg.tc_enabled = False; sg.fused_layer_enabled = False
    try:
        with layers.first_order_only(False):
            for _ in range(2):
                tr.step(real, gen)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            K = 3
            e0.record()
            for _ in range(K):
                tr.step(real, gen)
            e1.record(); torch.cuda.synchronize()
        msb = e0.elapsed_time(e1) / K
    finally:
        cg.tc_enabled = True; sg.fused_layer_enabled = True
        tr.micro_batch, tr.batch_size = old_mb, old_bs
    return dict(value=Bb / (msb * 1e-3), unit='images/s', ms_per_step=msb, batch=Bb, steps=K,
                kind='cuDNN fp32 (TF32 off) convolutions, unfused layer composition; same modules / weights / ray-march kernel',
                note='first_order_only(False) forces the unfused D path for every phase; steps 2..4 of this leg contain no Dreg phase')


def run_ginfer(args, rank, world, local):
    """BASELINE configs[4]: G-only inference (the FID2k path, metric_utils.py:303-319): G(z, c, camera) in eval mode, full-frame
    256x256 render (65 536 rays / image, 48+48 samples / ray), images converted to uint8 on the device and read back."""
    gp = importlib.import_module('3dgp_b200')
    cfgm = importlib.import_module('3dgp_b200.config'); dn = importlib.import_module('3dgp_b200.dnnlib')
    dev = torch.device('cuda', local)
    inf = importlib.import_module('3dgp_b200.training.inference')
    B = args.batch_gpu or 64                 # BASELINE configs[4]: batch 64
    cfg = cfgm.make_config(batch_size=B * world)
    torch.manual_seed(99 + rank)
    G, _ = cfgm.build_networks(cfg, dev)
    G.eval().requires_grad_(False)
    host = synthetic_batch(cfg, B, dev, seed=rank)
    keys = ('z', 'c', 'angles', 'fov', 'radius', 'look_at')
    d = {k: host[k].to(dev) for k in keys}
    cam = dn.TensorGroup(angles=d['angles'], fov=d['fov'], radius=d['radius'], look_at=d['look_at'])

    gg = inf.GraphedGenerator(G, B, noise_mode='const') if args.graph else None       # CUDA-graph replay of the whole generator call

    def step():
        if gg is not None:
            return gg(d['z'], d['c'], cam)
        return inf.generate_uint8(G, d['z'], d['c'], cam, noise_mode='const')        # metric_utils.py:306-313 on the fused path

    c0 = gp._lib.launch_count
    ms, clocks = timed_region(step, args.steps, args.warmup, world)
    launches = (gp._lib.launch_count - c0) // (args.steps + max(args.warmup, 3))
    out_h = torch.empty([B, 3, cfg.dataset.resolution, cfg.dataset.resolution], dtype=torch.uint8, pin_memory=True)

    def step_e2e():
        dd = {k: host[k].to(dev, non_blocking=True) for k in keys}
        cm = dn.TensorGroup(angles=dd['angles'], fov=dd['fov'], radius=dd['radius'], look_at=dd['look_at'])
        out_h.copy_(gg(dd['z'], dd['c'], cm) if gg is not None else inf.generate_uint8(G, dd['z'], dd['c'], cm, noise_mode='const'), non_blocking=True)

    ms_e2e, _ = timed_region(step_e2e, args.steps, 1, world)
    fg, _ = conv_flops_per_image(cfg)
    peaks = measured_peaks()
    R_ = cfg.dataset.resolution ** 2
    alg = raymarch_algorithmic_bytes(B, R_, cfg.model.generator.num_ray_steps, 512, 32)
    return dict(metric='G-only inference images/s at 256x256', value=world * B / (ms * 1e-3), unit='images/s', ms_per_step=ms,
                dtype='fp32 (bf16x3 tensor-core convs, 3xTF32 tri-plane MLP)',
                config=dict(workload='ginfer (BASELINE configs[4]: G-only inference, 256x256 full-frame render)', batch_per_gpu=B, cuda_graph=bool(args.graph),
                            rays_per_image=R_, l2='tri-planes of the batch (%.1f GB) larger than the 126 MB L2' % (B * 100.66e6 / 1e9),
                            parallelism=f'replicas x{world} (no collective)'),
                roofline=dict(bound='tensor', achieved=B * fg / (ms * 1e-3) / 1e12, peak=peaks['bf16_tflops_sustained'], unit='TFLOP/s',
                              frac=B * fg / (ms * 1e-3) / 1e12 / peaks['bf16_tflops_sustained'], traffic=None, kernel='conv_nhwc_bf16_kernel (whole forward)',
                              peak_source=peaks['source'], note='decoder conv FLOPs (x1, not x3) of the batch / step time; ray-march bytes/launch = %d' % alg),
                e2e=dict(value=world * B / (ms_e2e * 1e-3), unit='images/s', h2d_bytes_per_step=int(sum(host[k].numel() * host[k].element_size() for k in keys)),
                         d2h_bytes_per_step=int(out_h.numel())),
                gpu_launches=int(launches * args.steps), clocks=clocks)


# ---------------------------------------------------------------------------------------------------------------------
# workload: op-level microbench (BASELINE configs[3]: upsample2d / upfirdn2d / bias_act / filtered_lrelu against the HBM roofline)
def run_ops(args, rank, world, local):
    """Plugin ops on their own: every case reports algorithmic bytes (input + output once) / time.  The headline
    value is the byte-weighted aggregate GB/s over all cases; per-case numbers are in config.cases."""
    up = importlib.import_module('3dgp_b200.torch_utils.ops.upfirdn2d')
    ba = importlib.import_module('3dgp_b200.torch_utils.ops.bias_act')
    fl = importlib.import_module('3dgp_b200.torch_utils.ops.filtered_lrelu')
    gp = importlib.import_module('3dgp_b200')
    dev = torch.device('cuda', local)
    torch.manual_seed(5 + rank)
    f4 = up.setup_filter([1, 3, 3, 1], device=dev)
    f12 = up.setup_filter(np.kaiser(12, 8.0), device=dev)
    B = args.batch_gpu or 16
    cases = []

    def add(name, fn, x, note=''):
        y = fn(x)
        cases.append(dict(name=name, fn=fn, x=x, bytes=x.numel() * x.element_size() + y.numel() * y.element_size(), note=note))
        del y

    cl = lambda t: t.contiguous(memory_format=torch.channels_last)
    # the FIR that follows G's up-sampling conv (conv2d_resample.py:119-126): [B,128,513,513] -> 512^2, fp32, channel-minor
    add('upfirdn2d pad1 gain4 f32 cl [B,128,513,513]', lambda t: up.upfirdn2d(t, f4, padding=1, gain=4), cl(torch.randn(B, 128, 513, 513, device=dev)))
    add('upfirdn2d pad1 gain4 f32 cl [B,512,129,129]', lambda t: up.upfirdn2d(t, f4, padding=1, gain=4), cl(torch.randn(B, 512, 129, 129, device=dev)))
    # skip-image up-sampling of the tri-plane decoder (networks_stylegan2.py:268): 96 channels
    add('upsample2d x2 f32 cl [B,96,256,256]', lambda t: up.upsample2d(t, f4), cl(torch.randn(B, 96, 256, 256, device=dev)))
    # config 4 proper: fp16 NCHW, r: 64 -> 128 -> 256, C in {512, 256, 128}
    for C, r in ((512, 64), (256, 128), (128, 256)):
        add(f'upsample2d x2 f16 nchw [B,{C},{r},{r}]', lambda t: up.upsample2d(t, f4), torch.randn(B, C, r, r, device=dev).half())
        add(f'downsample2d x2 f16 nchw [B,{C},{2 * r},{2 * r}]', lambda t: up.downsample2d(t, f4), torch.randn(B, C, 2 * r, 2 * r, device=dev).half())
    bias = torch.randn(512, device=dev)
    add('bias_act lrelu f16 nchw [B,512,128,128]', lambda t: ba.bias_act(t, bias.half(), act='lrelu', clamp=256), torch.randn(B, 512, 128, 128, device=dev).half())
    add('bias_act lrelu f32 cl [B,128,512,512]', lambda t: ba.bias_act(t, bias[:128], act='lrelu'), cl(torch.randn(B, 128, 512, 512, device=dev)))
    add('filtered_lrelu up2 down2 12-tap f16 [B,256,128,128]', lambda t: fl.filtered_lrelu(t, fu=f12, fd=f12, b=bias[:256].half(), up=2, down=2, padding=[11, 10, 11, 10]),
        torch.randn(B, 256, 128, 128, device=dev).half(), note='fused single kernel (separable filters): csrc/filtered_lrelu.cu')
    peaks = measured_peaks()
    c0 = gp._lib.launch_count
    W_, K_ = max(args.warmup, 3), args.steps
    tot_b = tot_ms = 0.0
    rows = []
    clocks_mon = ClockSampler(local); clocks_mon.start()
    for c_ in cases:
        for _ in range(W_):
            c_['fn'](c_['x'])
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(K_):
            c_['fn'](c_['x'])
        ev[1].record(); torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / K_
        gbs = c_['bytes'] / (ms * 1e-3) / 1e9
        rows.append(dict(op=c_['name'], ms=round(ms, 4), gbs=round(gbs, 1), frac=round(gbs / peaks['hbm_gbs'], 3), **({'note': c_['note']} if c_['note'] else {})))
        tot_b += c_['bytes']; tot_ms += ms
    launches = gp._lib.launch_count - c0
    clocks = clocks_mon.stop()
    agg = tot_b / (tot_ms * 1e-3) / 1e9
    return dict(metric='plugin ops aggregate GB/s (algorithmic bytes of all cases / total time)', value=agg * world, unit='GB/s', ms_per_step=tot_ms, dtype='f16 / f32',
                config=dict(workload='ops (BASELINE configs[3]: op-level upfirdn2d / bias_act / filtered_lrelu)', batch_per_gpu=B,
                            l2='every case moves >= 130 MB (> the 126 MB L2)', cases=rows, parallelism=f'replicas x{world}'),
                roofline=dict(bound='hbm', achieved=agg, peak=peaks['hbm_gbs'], unit='GB/s', frac=agg / peaks['hbm_gbs'], traffic=None,
                              kernel='upfirdn2d / bias_act kernels (byte-weighted aggregate)', peak_source=peaks['source']),
                e2e=dict(value=agg * world, unit='GB/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0, note='device-resident microbench: no host leg'),
                gpu_launches=int(launches), clocks=clocks)


class CpuStep:
    """Reference algorithm on the host cores (oracle port: oracle/restated.py with DIFFERENTIABLE = True, driven by oracle/train_step.py) for the
    training-step metric.  One call of `run()` EXECUTES one optimisation step -- Gmain and Dmain forward + backward + Adam, Dreg (R1 double backward)
    on the reference's schedule (every 16th step) -- at the BASELINE widths on a bounded sample of `sample_batch` images (default 1; the minibatch-std
    group is then min(4, N) = 1 as in networks_discriminator.py:111).  Nothing is extrapolated."""

    def __init__(self, small=False, sample_batch=1):
        from oracle import ref_harness as rh, train_step as ts
        self.threads = min(os.cpu_count(), 32)      # beyond ~32 threads ATen's CPU convs stop scaling on these hosts
        torch.set_num_threads(self.threads)
        kw = dict(cmax=64, cbase=4096, tri_res=128, patch_res=32, img_resolution=128, c_dim=10, w_dim=128, z_dim=128, num_ray_steps=12) if small else {}
        Gc, Dc, m = rh.make_cfg(**kw)
        sdG, sdD = ts.random_state_dicts(Gc, Dc, m, seed=0)
        Bs = sample_batch
        r1_gamma = 0.0002 * (m['img_resolution'] ** 2) / 32
        self.tr = ts.CpuTrainer(sdG, sdD, Gc, Dc, m, d_reg_interval=16, r1_gamma=r1_gamma)
        g = torch.Generator().manual_seed(2)
        res = m['img_resolution']
        self.batch = dict(real_img=torch.randint(0, 256, (Bs, 3, res, res), generator=g).float() / 127.5 - 1,
                          real_depth=torch.randint(0, 65536, (Bs, 1, res, res), generator=g).float() / 65536 * 2 - 1,
                          c=torch.nn.functional.one_hot(torch.arange(Bs) % Gc['c_dim'], Gc['c_dim']).float(),
                          embs=torch.randn(Bs, m['embedding_dim'], generator=g), z=torch.randn(Bs, Gc['z_dim'], generator=g),
                          cam=dict(angles=torch.tensor([[0.3, 1.4, 0.0]]).repeat(Bs, 1), fov=torch.full((Bs,), 25.0), radius=torch.ones(Bs),
                                   look_at=torch.tensor([[0.1, 1.5, 0.1]]).repeat(Bs, 1)))
        self.Bs = Bs

    def run(self):
        t0 = time.perf_counter()
        st = self.tr.step(**self.batch)
        return time.perf_counter() - t0, st


def cpu_train_step(small=False, steps=1, warmup=0, sample_batch=1):
    """`steps` executed CPU optimisation steps after `warmup` untimed ones; returns the cpu_baseline object and the mean seconds per step."""
    cs = CpuStep(small=small, sample_batch=sample_batch)
    for _ in range(warmup):
        cs.run()
    ts_, had_reg = [], 0
    for _ in range(steps):
        dt, st = cs.run()
        ts_.append(dt); had_reg += int('Loss/D/r1_penalty' in st)
    mean_s = float(np.mean(ts_))
    return dict(value=cs.Bs / mean_s, unit='images/s', cores=cs.threads, kind='port',
                sample=f'{steps} executed optimisation step(s) (Gmain + Dmain forward/backward/Adam; {had_reg} with the lazy R1 phase) of {cs.Bs} image(s) at the '
                       f'{"BASELINE widths (cmax=1024, 512^2 tri-planes, 64^2 patch, 48+48 samples/ray)" if not small else "SMALL debug widths"}, {mean_s:.2f} s/step, after {warmup} untimed step(s)'), mean_s


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=os.environ.get('GP3D_BENCH_WORKLOAD', 'train_step'), choices=['train_step', 'raymarch', 'ginfer', 'ops'])
    ap.add_argument('--batch-gpu', type=int, default=0)
    ap.add_argument('--mlp-mode', type=int, default=2, help='tri-plane MLP arithmetic: 0 fp32 SIMT (v1 kernel), 1 TF32 mma, 2 3xTF32 mma (default)')
    ap.add_argument('--planes-fp16', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gpu-baseline', action='store_true')
    ap.add_argument('--no-ginfer', action='store_true', help='skip the G-only inference leg (BASELINE configs[4]) of the default run')
    ap.add_argument('--micro-batch', type=int, default=0)
    ap.add_argument('--small', action='store_true')
    ap.add_argument('--graph', action='store_true', help='ginfer: replay the generator call as a CUDA graph (training/inference.py::GraphedGenerator)')
    ap.add_argument('--overlap', action='store_true', help='experimental: 64 MB gradient buckets all-reduced during the final backward of each phase (validated on 2 ranks only)')
    ap.add_argument('--cpu-sample-batch', type=int, default=1, help='images per executed CPU step of the reference arm / cpu_baseline leg')
    args = ap.parse_args()
    rank, world, local = dist_info()

    if args.impl == 'reference':
        if rank != 0:
            return
        if args.workload == 'train_step':
            B = args.batch_gpu or 32
            cb, mean_s = cpu_train_step(small=args.small, steps=args.steps, warmup=args.warmup, sample_batch=args.cpu_sample_batch)
            metric = 'G+D training-step images/s at 256x256'
            cfg_ = train_step_config(B, args.micro_batch or min(B, 32), args.gpus, args.small)
            ms_step = mean_s * 1e3
        else:
            cb = cpu_raymarch(sample_rays=4096, repeats=max(args.steps - 1, 1))
            metric = 'ray-march images/s (64x64 rays, 48+48 samples/ray, 32-ch 512^2 tri-planes)'
            cfg_ = dict(workload='raymarch (BASELINE configs[2])')
            ms_step = 1e3 / cb['value']
        line = dict(impl='reference', metric=metric, value=cb['value'], unit='images/s',
                    n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step, higher_is_better=True, scaling='weak',
                    vs_baseline=None, dtype='f32', data='synthetic', config=cfg_,
                    cpu_baseline=cb, e2e=dict(value=cb['value'], unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl')
    res = {'train_step': run_train_step, 'raymarch': run_raymarch, 'ginfer': run_ginfer, 'ops': run_ops}[args.workload](args, rank, world, local)
    if rank == 0:
        line = dict(metric=res['metric'], value=res['value'], unit=res['unit'], n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=res['ms_per_step'], higher_is_better=True, scaling='weak', vs_baseline=None, dtype=res['dtype'], data='synthetic',
                    config=res['config'], roofline=res['roofline'], e2e=res['e2e'], gpu_launches=res['gpu_launches'], clocks=res['clocks'])
        for k in ('details', 'roofline_step_tensor', 'roofline_raymarch', 'forward_backward', 'gpu_baseline', 'ginfer'):
            if k in res:
                line[k] = res[k]
        if world == 1 and not args.no_cpu_baseline:
            if args.workload == 'train_step':
                cb = _side_leg('cpu_baseline', cpu_train_step, small=args.small, steps=1, warmup=0, sample_batch=args.cpu_sample_batch)
                line['cpu_baseline'] = cb if isinstance(cb, dict) else cb[0]
            elif args.workload == 'raymarch':
                line['cpu_baseline'] = _side_leg('cpu_baseline', cpu_raymarch, sample_rays=4096, repeats=2)
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
