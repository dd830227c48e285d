"""GPU suite: fused all-reduce epilogue + Adam + G_ema kernel (csrc/optim.cu) against torch.optim.Adam, the optimiser the
reference builds at training_loop.py:190-205 (third-party arithmetic: torch/optim/adam.py, foreach path)."""
import copy
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


def _nets(seed=0):
    torch.manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.Linear(37, 129), torch.nn.Linear(129, 1031), torch.nn.Linear(1031, 5, bias=False)).cuda()
    return net, copy.deepcopy(net), copy.deepcopy(net).requires_grad_(False), None


@pytest.mark.parametrize('betas', [(0.0, 0.99), (0.9, 0.999), (0.0, 0.99 ** (16 / 17))])
def test_flat_adam_tracks_torch_adam_with_nan_to_num_and_ema(betas):
    stepm = importlib.import_module('3dgp_b200.training.step')
    net, ref, ema, _ = _nets()
    ema_ref = copy.deepcopy(ema)
    opt = stepm.FlatAdam(net, lr=0.002, betas=betas, eps=1e-8, ema_module=ema)
    ropt = torch.optim.Adam(ref.parameters(), lr=0.002, betas=betas, eps=1e-8)
    g = torch.Generator(device='cuda').manual_seed(1)
    for it in range(6):
        opt.zero_grad(); ropt.zero_grad(set_to_none=True)
        x = torch.randn(8, 37, device='cuda', generator=g)
        skip_last = it in (2, 3)                       # the last layer gets no gradient in two of the phases
        h = net[1](net[0](x))
        loss = (h.square().mean() if skip_last else net[2](h).square().mean()) * (3.0 ** it)
        loss.backward()
        if it == 4:                                    # non-finite gradients: nan -> 0, +-inf -> +-1e5 (training_loop.py:341)
            net[0].weight.grad[0, 0] = float('nan'); net[0].weight.grad[0, 1] = float('inf'); net[1].bias.grad[3] = -float('inf')
        # the same gradients go to torch.optim.Adam, so that only the optimiser arithmetic is compared
        for i, (p, q) in enumerate(zip(net.parameters(), ref.parameters())):
            q.grad = torch.nan_to_num(p.grad.detach().clone(), nan=0, posinf=1e5, neginf=-1e5) if i in opt.active else None
        opt.step(ema_beta=0.75 if it % 2 else 0.25)
        ropt.step()
        with torch.no_grad():
            for pe, p in zip(ema_ref.parameters(), ref.parameters()):
                pe.copy_(p.lerp(pe, 0.75 if it % 2 else 0.25))
        for (n, a), b in zip(net.named_parameters(), ref.parameters()):
            assert torch.isfinite(a).all()
            err = (a - b).abs().max().item()
            assert err <= 2e-6 * max(1.0, b.abs().max().item()), (it, n, err)
        for a, b in zip(ema.parameters(), ema_ref.parameters()):
            assert (a - b).abs().max().item() <= 2e-6 * max(1.0, b.abs().max().item())
    assert opt.steps[-1] == 4 and opt.steps[0] == 6    # skipped phases do not advance that tensor's step count
    sd = opt.state_dict()
    rs = ropt.state_dict()['state']
    for i in range(len(opt.params)):
        assert sd['state'][i]['step'] == int(rs[i]['step'])
        assert torch.allclose(sd['state'][i]['exp_avg_sq'], rs[i]['exp_avg_sq'], rtol=2e-6, atol=1e-30)


def test_flat_adam_parameters_are_views_of_one_storage_and_grads_accumulate_in_place():
    stepm = importlib.import_module('3dgp_b200.training.step')
    net, _, _, _ = _nets()
    opt = stepm.FlatAdam(net, lr=1e-3, betas=(0.0, 0.99))
    base = opt.flat_p.data_ptr()
    for p, o in zip(opt.params, opt.offsets):
        assert p.data_ptr() == base + 4 * o and o % 1024 == 0
    opt.zero_grad()
    x = torch.randn(4, 37, device='cuda')
    for _ in range(2):                                  # two micro-batches accumulate into the flat gradient
        net(x).sum().backward()
    for p, gv in zip(opt.params, opt.grad_views):
        assert p.grad.data_ptr() == gv.data_ptr()
    ref = copy.deepcopy(net)
    for p in ref.parameters():
        p.grad = None
    ref(x).sum().backward()
    for p, q in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, 2 * q.grad, rtol=1e-5, atol=1e-6)
    assert len(opt.active) == len(opt.params)


def test_adam_kernel_rejects_unaligned_storage():
    _lib = importlib.import_module('3dgp_b200._lib')
    t = torch.zeros(1000, device='cuda')
    rc = _lib.lib().gp3d_adam_ema_step(t.data_ptr(), t.data_ptr(), None, t.data_ptr(), None, 1000, 1.0, 1e5, -1e5, 0.0, 0.99, 1.0, 0.01, 1e-8, 1e-3, 1.0, 0.0, None, None, None)
    assert rc != 0 and b'1024' in _lib.lib().gp3d_last_error()
