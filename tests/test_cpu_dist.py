"""CPU suite: the N>1 host logic (flattened gradient all-reduce of training_loop.py:335-344) with gloo, world_size 2."""
import importlib
import os
import socket

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
    out = [p.grad.clone() for p in net.parameters()]
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
    assert n0 == n1 == 5 * 7 + 7 + 7 * 3 + 3
    for a, b in zip(g0, g1):
        assert torch.equal(a, b)                     # every rank ends with the same averaged gradient
    # mean of (i+1)*1 and (i+1)*2 = 1.5*(i+1); nan -> 0 ; +inf -> 1e5 (nan_to_num semantics of training_loop.py:341)
    assert torch.allclose(g0[0][1:], torch.full_like(g0[0][1:], 1.5))
    assert g0[0][0, 0].item() == 0.0
    assert g0[1][0].item() == 1e5
    assert torch.allclose(g0[2], torch.full_like(g0[2], 4.5))
