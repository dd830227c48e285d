"""CPU suite: the N>1 host logic (flattened gradient all-reduce of training_loop.py:335-344) with gloo, world_size 2 (and 4 for the bucketed form)."""
import importlib
import os
import socket

import pytest

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    stepm = importlib.import_module('3dgp_b200.training.step')
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    list(net.parameters())[0].grad[0, 0] = float('nan') if rank == 0 else 1.0
    list(net.parameters())[1].grad[0] = float('inf') if rank == 1 else 1.0
    n = stepm.allreduce_gradients(list(net.parameters()), world)
    out = [p.grad.clone().numpy() for p in net.parameters()]      # by value: a tensor would travel as a shared-memory handle that dies with this process
    q.put((rank, n, out))
    dist.destroy_process_group()


def test_flattened_gradient_allreduce_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    (_, n0, g0), (_, n1, g1) = res
    g0, g1 = [torch.from_numpy(a) for a in g0], [torch.from_numpy(a) for a in g1]
    assert n0 == n1 == 5 * 7 + 7 + 7 * 3 + 3
    for a, b in zip(g0, g1):
        assert torch.equal(a, b)                     # every rank ends with the same averaged gradient
    # mean of (i+1)*1 and (i+1)*2 = 1.5*(i+1); nan -> 0 ; +inf -> 1e5 (nan_to_num semantics of training_loop.py:341)
    assert torch.allclose(g0[0][1:], torch.full_like(g0[0][1:], 1.5))
    assert g0[0][0, 0].item() == 0.0
    assert g0[1][0].item() == 1e5
    assert torch.allclose(g0[2], torch.full_like(g0[2], 4.5))


def _bucket_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    stepm = importlib.import_module('3dgp_b200.training.step')
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    params = list(net.parameters())
    params[-1].requires_grad_(False)                 # a parameter that never receives a gradient: its bucket can only be reduced by finish()
    offsets, off = [], 0
    for p in params:
        offsets.append(off); off += -(-p.numel() // 8) * 8
    flat = torch.zeros(off)
    for p, o in zip(params, offsets):
        p.grad = flat[o:o + p.numel()].view(p.shape)
    bk = stepm.GradBuckets(flat, offsets, [p.numel() for p in params], bucket_elems=64)
    for i, p in enumerate(params):
        if p.requires_grad:
            p.register_post_accumulate_grad_hook(lambda _p, i=i: bk.ready(i))
    x = torch.randn(5, 6, generator=torch.Generator().manual_seed(10 + rank))
    net(x).square().sum().backward()                 # not armed: a non-final backward of the phase accumulates without any transfer
    assert bk.launched_async == 0
    bk.arm(world)                                    # first armed pass: nothing known about the graph, the frozen tail parameter blocks bucket 0
    net(x).square().sum().backward()
    first = bk.launched_async
    bk.finish()
    fired = set(bk.fired)
    flat.zero_()
    net(x).square().sum().backward()                 # this rank's own gradient, for the check
    local = flat.clone()
    flat.zero_()
    bk.arm(world, expected=fired)                    # steady state: buckets stream out as they complete
    net(x).square().sum().backward()
    launched_during_backward = bk.launched_async
    bk.finish()
    q.put((rank, len(bk.bounds), (first, launched_during_backward), local.numpy(), flat.clone().numpy()))
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4])
def test_bucketed_gradient_allreduce_overlaps_and_matches_one_shot_gloo(world):
    """GradBuckets (training/step.py): buckets cover the flat buffer exactly once, launch in index order while the final backward is still producing
    gradients, and the result equals ONE all-reduce of the whole buffer (training_loop.py:335-344).  World 4 is the rank count at which the NCCL run
    of the overlapped form stalled on the B200 box: the host logic -- same collective sequence on every rank -- holds at 4 ranks."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    nb, launches = res[0][1], res[0][2]
    assert nb >= 3 and all(r[1] == nb and r[2] == launches for r in res)
    assert launches[0] == 0 and launches[1] == nb    # pass 1 learns which parameters fire; pass 2 sends every bucket while the backward runs
    reduced = [torch.from_numpy(r[4]) for r in res]
    assert all(torch.equal(reduced[0], t) for t in reduced[1:])
    want = torch.stack([torch.from_numpy(r[3]).double() for r in res]).sum(0)
    assert torch.allclose(reduced[0].double(), want, rtol=1e-6, atol=1e-6)


class _LateTail(torch.nn.Module):
    """Like the Generator (networks_epigraf.py: `synthesis` registered before `mapping`): the parameters at the END of the flat buffer belong to the layer
    that runs FIRST, so their gradients arrive last."""
    def __init__(self):
        super().__init__()
        self.body = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
        self.first = torch.nn.Linear(6, 6)

    def forward(self, x):
        return self.body(self.first(x))


def _late_tail_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    stepm = importlib.import_module('3dgp_b200.training.step')
    torch.manual_seed(0)
    net = _LateTail()
    params = list(net.parameters())
    offsets, off = [], 0
    for p in params:
        offsets.append(off); off += -(-p.numel() // 8) * 8
    flat = torch.zeros(off)
    for p, o in zip(params, offsets):
        p.grad = flat[o:o + p.numel()].view(p.shape)
    bk = stepm.GradBuckets(flat, offsets, [p.numel() for p in params], bucket_elems=64)
    for i, p in enumerate(params):
        p.register_post_accumulate_grad_hook(lambda _p, i=i: bk.ready(i))
    x = torch.randn(5, 6, generator=torch.Generator().manual_seed(20 + rank))
    net(x).square().sum().backward()
    local = flat.clone(); flat.zero_()
    bk.arm(world)                                    # index order: bucket 0 (the tail = `first`) completes last and holds every other bucket back
    net(x).square().sum().backward()
    log_index = list(bk.launch_log)
    bk.finish()
    reduced_index = flat.clone(); flat.zero_()
    order = bk.completion_order()
    bk.arm(world, expected=set(bk.fired), order=order)   # previous completion order: transfers start as soon as the body's last layers are done
    net(x).square().sum().backward()
    log_order = list(bk.launch_log)
    bk.finish()
    q.put((rank, len(bk.bounds), order, log_index, log_order, local.numpy(), reduced_index.numpy(), flat.clone().numpy()))
    dist.destroy_process_group()


def test_bucket_launch_order_follows_the_previous_completion_order_gloo_world2():
    """With the first layer's parameters at the tail of the buffer (the Generator's layout), index-ordered launching starts every transfer only after the
    LAST gradient; launching in the previous pass's completion order starts the first transfer after the first bucket's gradients -- same sums either way,
    same sequence on both ranks."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_late_tail_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    (_, nb, order, log_i, log_o, loc0, ri0, ro0), (_, nb1, order1, log_i1, log_o1, loc1, ri1, ro1) = res
    n_params = 8
    assert nb == nb1 >= 3 and order == order1 and log_i == log_i1 and log_o == log_o1          # identical collective sequence on both ranks
    assert order[-1] == 0 and order != list(range(nb))                                          # the tail bucket completes last
    assert [b for b, _ in log_i] == list(range(nb)) and all(k == n_params for _, k in log_i)    # index order: everything waits for the last gradient
    assert [b for b, _ in log_o] == order and log_o[0][1] <= n_params // 2                      # completion order: the first transfer starts early
    want = torch.from_numpy(loc0) + torch.from_numpy(loc1)
    for r in (ri0, ri1, ro0, ro1):
        assert torch.equal(torch.from_numpy(r), want)


def _trainer_worker(rank, world, port, q):
    """One optimisation step of the REAL Generator / Discriminator / loss on each rank (CPU modules on the emulated C ABI, tests/abi_emulator.py), gloo
    carrying the per-phase gradient all-reduce: ranks start from different initialisations and see different data."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import json
    import sys
    import numpy as np
    import pytest
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, 'tests'))
    import abi_emulator as emu
    from oracle import cases
    emu.install(pytest.MonkeyPatch())
    torch.set_num_threads(4)
    cfgm = importlib.import_module('3dgp_b200.config'); dn = importlib.import_module('3dgp_b200.dnnlib')
    lossm = importlib.import_module('3dgp_b200.training.loss'); stepm = importlib.import_module('3dgp_b200.training.step')
    meta = json.load(open(os.path.join(root, 'tests', 'golden', 'networks_meta.json')))
    kw = dict(meta['net_kwargs']); kw.pop('learn_camera_dist', None)
    cfg = cfgm.make_config(**kw, kd_weight=1.0, batch_size=4 * world)
    torch.manual_seed(100 + rank); np.random.seed(100 + rank)               # rank-specific initialisation: Trainer must broadcast rank 0's (training_loop.py:176-179)
    G, D = cfgm.build_networks(cfg, 'cpu', fp32_D=True)
    probe = lambda m: float(sum(p.detach().double().sum() for p in m.parameters()))
    before = (probe(G), probe(D))
    loss = lossm.StyleGAN2Loss(cfg, 'cpu', G, D, r1_gamma=1.0)
    tr = stepm.Trainer(G, D, loss, cfg, rank=rank, world_size=world, D_reg_interval=16, batch_size=4 * world)
    synced = (probe(G), probe(D))
    t = {k: torch.from_numpy(v) for k, v in cases.net_inputs(meta['net_kwargs']).items()}
    B, res = t['z'].shape[0], kw['img_resolution']
    g = torch.Generator().manual_seed(7 + rank)                              # rank-specific data
    real = dn.EasyDict(img=torch.rand(B, 3, res, res, generator=g) * 2 - 1, depth=torch.rand(B, 1, res, res, generator=g) * 2 - 1, c=t['c'],
                       embs=torch.randn(B, kw['embedding_dim'], generator=g), camera_angles=t['angles'])
    gen = dn.EasyDict(z=torch.randn(B, kw['z_dim'], generator=g), c=t['c'],
                      camera_params=dn.TensorGroup(angles=t['angles'], fov=t['fov'], radius=t['radius'], look_at=t['look_at']))
    stats = tr.step(real, gen)
    finite = all(bool(torch.isfinite(torch.as_tensor(v)).all()) for v in stats.values())
    q.put((rank, before, synced, (probe(G), probe(D)), probe(tr.G_ema), finite))
    dist.destroy_process_group()


def test_real_networks_stay_replicated_over_one_step_gloo_world2():
    """The data-parallel invariant of training_loop.py:176-179 + :335-344 on the real modules: after the start-up broadcast and one iteration on
    different data, every rank holds the same G, D and G_ema."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_trainer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    (_, b0, s0, a0, e0, f0), (_, b1, s1, a1, e1, f1) = res
    assert f0 and f1
    assert b0 != b1                                   # different initialisations ...
    assert s0 == s1 == b0                             # ... replaced by rank 0's parameters
    assert a0 == a1 and a0 != s0                      # one step on different data: identical parameters on both ranks, and they moved
    assert e0 == e1
