#!/usr/bin/env python
"""bench.py -- one JSON line per run (driver contract, see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train_step|raymarch|ginfer] [--impl ours|reference]

Workloads (BASELINE.json `configs`):
  train_step : configs[1]  ImageNet-256 G+D training step (Gmain + Dmain phases), cmax=1024, 48 samples/ray,
               patch 64x64, batch/GPU from --batch-gpu, synthetic data, random-init weights.  metric = images/s.
  raymarch   : configs[2]  fused ray-march microbench, 64x64 rays, 48+48 samples/ray, 32-ch 512^2 tri-planes, batch 16.
  ginfer     : configs[4]  G-only inference at 256x256.
`--impl reference` times the reference's CPU algorithm (the oracle port, torch-CPU/numpy, all host threads) on a
bounded sample of the same workload; under torchrun only rank 0 runs it.

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
    """Synthetic tri-planes N(0,1) stored channel-minor (the layout the tri-plane decoder emits), full-frame 64x64 rays
    from cameras ~ configs/camera/{base,uniform}.yaml, random-init MLP (layers.py:36: randn weights, zero bias)."""
    from oracle import restated as R
    g = torch.Generator(device='cpu').manual_seed(seed)
    P, C = RM['P'], RM['C']
    planes = torch.randn([B, P, P, 3 * C], device=device, generator=torch.Generator(device=device).manual_seed(seed))
    planes = planes.permute(0, 3, 1, 2).view(B, 3, C, P, P)
    yaw = torch.rand(B, generator=g) * 3.14 - 1.57
    pitch = torch.rand(B, generator=g) * (2.35619449 - 0.785398163) + 0.785398163
    angles = torch.stack([yaw, pitch, torch.zeros(B)], 1)
    fov = torch.rand(B, generator=g) * 35 + 10
    look = torch.stack([torch.rand(B, generator=g) * 6.28 - 3.14, torch.acos(1 - 2 * torch.rand(B, generator=g).clamp(1e-5, 1 - 1e-5)), torch.rand(B, generator=g) * 0.2], 1)
    c2w = R.compute_cam2world_matrix(angles, torch.ones(B), look)
    ro, rd = R.sample_rays(c2w, fov, (RM['res'], RM['res']))
    w1 = torch.randn(RM['H'], C, generator=g); w2 = torch.randn(4, RM['H'], generator=g)
    return dict(planes=planes, ray_o=ro, ray_d=rd, w1=w1, b1=torch.zeros(RM['H']), w2=w2, b2=torch.zeros(4))


def raymarch_algorithmic_bytes(B, R, N, P, C, plane_bytes=4, injected_noise=False):
    """planes once + rays in (+ injected variates) + results out."""
    return B * 3 * C * P * P * plane_bytes + B * R * (24 + (2 * N * 4 if injected_noise else 0)) + B * R * 24


def run_raymarch(args, rank, world, local):
    rmod = importlib.import_module('3dgp_b200.torch_utils.ops.raymarch')
    dev = torch.device('cuda', local)
    B = args.batch_gpu or RM['B']
    inp = raymarch_inputs(B, dev, seed=rank)
    R_ = RM['res'] ** 2
    d = {k: (v.to(dev) if k != 'planes' else v) for k, v in inp.items()}
    kw = dict(num_steps=RM['N'], ray_start=RM['ray_start'], ray_end=RM['ray_end'], box_size=2 * RM['box_half'], mlp_mode=args.mlp_mode)
    launches = {'n': 0}
    pl = rmod.planes_channel_minor(d['planes'])
    if args.planes_fp16:
        pl = pl.half()

    def step():
        out = rmod.render_rays(pl, d['w1'], d['b1'], d['w2'], d['b2'], d['ray_o'], d['ray_d'], seed=launches['n'], **kw)
        launches['n'] += 1
        return out

    ms, clocks = timed_region(step, args.steps, args.warmup, world)
    n0 = launches['n']
    # end-to-end: rays from pinned host memory every step, results read back to the host
    ro_h = inp['ray_o'].pin_memory(); rd_h = inp['ray_d'].pin_memory()
    out_h = torch.empty([B, R_, 5], pin_memory=True)

    def step_e2e():
        ro = ro_h.to(dev, non_blocking=True); rd = rd_h.to(dev, non_blocking=True)
        rgb, depth, wsum, _ = rmod.render_rays(pl, d['w1'], d['b1'], d['w2'], d['b2'], ro, rd, seed=launches['n'], **kw)
        launches['n'] += 1
        out_h.copy_(torch.cat([rgb, depth, wsum], dim=-1), non_blocking=True)

    ms_e2e, _ = timed_region(step_e2e, args.steps, args.warmup, world)
    pb = 2 if args.planes_fp16 else 4
    alg = raymarch_algorithmic_bytes(B, R_, RM['N'], RM['P'], RM['C'], plane_bytes=pb)
    peaks = measured_peaks()
    ach = alg / (ms * 1e-3) / 1e9
    res = dict(
        metric='ray-march images/s (64x64 rays, 48+48 samples/ray, 32-ch 512^2 tri-planes)', value=world * B / (ms * 1e-3), unit='images/s',
        ms_per_step=ms, dtype='f32' if not args.planes_fp16 else 'f32 math / f16 planes',
        config=dict(workload='raymarch (BASELINE configs[2])', batch_per_gpu=B, rays=R_, samples_per_ray=2 * RM['N'], plane_res=RM['P'],
                    l2='inputs (%.2f GB of planes per GPU) larger than the 126 MB L2' % (B * 3 * RM['C'] * RM['P'] ** 2 * pb / 1e9),
                    rng='in-kernel Philox', mlp_mode=args.mlp_mode, parallelism=f'replicas x{world} (render does not shard)'),
        roofline=dict(bound='hbm', achieved=ach, peak=peaks['hbm_gbs'], unit='GB/s', frac=ach / peaks['hbm_gbs'], traffic=None,
                      kernel='raymarch_fwd_kernel', peak_source=peaks['source'], algorithmic_bytes_per_launch=alg),
        e2e=dict(value=world * B / (ms_e2e * 1e-3), unit='images/s', h2d_bytes_per_step=int(2 * B * R_ * 12), d2h_bytes_per_step=int(B * R_ * 20)),
        gpu_launches=args.steps, clocks=clocks)
    return res


def cpu_raymarch(sample_rays=1024, repeats=1):
    """Reference algorithm on the host cores (oracle port), bounded sample: one image, `sample_rays` rays."""
    from oracle import restated as R
    torch.set_num_threads(os.cpu_count())
    inp = raymarch_inputs(1, 'cpu', seed=0)
    planes = inp['planes'].contiguous()
    ro = inp['ray_o'][:, :sample_rays]; rd = inp['ray_d'][:, :sample_rays]
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
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=os.environ.get('GP3D_BENCH_WORKLOAD', 'raymarch'), choices=['train_step', 'raymarch', 'ginfer'])
    ap.add_argument('--batch-gpu', type=int, default=0)
    ap.add_argument('--mlp-mode', type=int, default=0)
    ap.add_argument('--planes-fp16', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank, world, local = dist_info()

    if args.impl == 'reference':
        if rank != 0:
            return
        cb = cpu_raymarch(sample_rays=2048, repeats=max(1, min(args.steps, 3)))
        line = dict(impl='reference', metric='ray-march images/s (64x64 rays, 48+48 samples/ray, 32-ch 512^2 tri-planes)', value=cb['value'], unit='images/s',
                    n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 / cb['value'], higher_is_better=True, scaling='weak',
                    vs_baseline=None, dtype='f32', data='synthetic', config=dict(workload='raymarch (BASELINE configs[2])'),
                    cpu_baseline=cb, e2e=dict(value=cb['value'], unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl')
    res = run_raymarch(args, rank, world, local)
    if rank == 0:
        line = dict(metric=res['metric'], value=res['value'], unit=res['unit'], n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=res['ms_per_step'], higher_is_better=True, scaling='weak', vs_baseline=None, dtype=res['dtype'], data='synthetic',
                    config=res['config'], roofline=res['roofline'], e2e=res['e2e'], gpu_launches=res['gpu_launches'], clocks=res['clocks'])
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_raymarch()
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
