"""GPU diagnostic: relative error (vs float64) of every convolution arithmetic in use -- ATen/cuDNN fp32 with TF32 off on the shapes that stay on ATen,
and the tcgen05 precision codes (ops.tc.operand_formats: 3 bf16x3, 2 x2w16, 16 fp16 class incl. the mixed bf16 x fp16 gradient forms, 1 bf16)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tc = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
cfgm = importlib.import_module('3dgp_b200.config')
cfgm.set_reference_numerics()
torch.manual_seed(0)
dev = 'cuda'


def rel(a, b):
    return ((a.double() - b).norm() / b.norm()).item(), ((a.double() - b).abs().max() / b.abs().max()).item()


print('--- ATen / cuDNN fp32 (allow_tf32 = False) vs float64')
for (cin, cout, k, r, n) in [(4, 64, 1, 32, 4), (129, 128, 3, 4, 4), (1, 64, 5, 32, 4), (64, 1, 1, 32, 4), (4, 256, 1, 64, 8)]:
    x = torch.randn(n, cin, r, r, device=dev, requires_grad=True); w = (torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5).requires_grad_(True)
    y = torch.nn.functional.conv2d(x, w, padding=k // 2)
    dy = torch.randn_like(y)
    gx, gw = torch.autograd.grad(y, [x, w], dy)
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    yd = torch.nn.functional.conv2d(xd, wd, padding=k // 2)
    gxd, gwd = torch.autograd.grad(yd, [xd, wd], dy.double())
    print(f'conv {cin}->{cout} k{k} @{r}: fwd l2/max {rel(y, yd)}  dgrad {rel(gx, gxd)}  wgrad {rel(gw, gwd)}')
    # channels-last matmul formulation for 1x1
    if k == 1:
        ym = torch.einsum('nchw,oc->nohw', x, w[:, :, 0, 0]); gxm = torch.einsum('nohw,oc->nchw', dy, w[:, :, 0, 0])
        print(f'   as einsum: fwd {rel(ym, yd)} dgrad {rel(gxm, gxd)}')

print('--- tcgen05 precision codes vs float64 (128 -> 128, 3x3, 32x32, N = 4)')
x = torch.randn(4, 128, 32, 32, device=dev); w = torch.randn(128, 128, 3, 3, device=dev) / (128 * 9) ** 0.5
yd = torch.nn.functional.conv2d(x.double(), w.double(), padding=1)
for terms, xg, wg, name in [(3, False, False, 'bf16x3'), (2, False, False, 'x2w16 (bf16 pair x fp16 weight)'), (16, False, False, 'f16 x f16'),
                            (16, True, False, 'bf16 x f16 (gradient x weight, mixed formats)'), (16, False, True, 'f16 x bf16 (mixed formats)'), (1, False, False, 'bf16 x bf16')]:
    try:
        y = tc.conv2d_forward(x, w, terms, x_is_grad=xg, w_is_grad=wg)
        torch.cuda.synchronize()
        print(f'terms {terms:2d} {name:48s} l2 / max rel err {rel(y, yd)}')
    except Exception as e:
        print(f'terms {terms} {name}: FAILED {e}')
print('--- weight gradient')
dy = torch.randn(4, 128, 32, 32, device=dev)
xd, wd = x.double(), w.double().requires_grad_(True)
gwd = torch.autograd.grad(torch.nn.functional.conv2d(xd, wd, padding=1), wd, dy.double())[0]
for terms, xg, name in [(3, False, 'bf16x3'), (16, False, 'bf16 dy x f16 x (mixed)'), (16, True, 'bf16 x bf16'), (1, False, 'bf16')]:
    try:
        gw = tc.conv_wgrad(dy, x, 3, 'conv', 1, 1, terms, x_is_grad=xg)
        torch.cuda.synchronize()
        print(f'terms {terms:2d} {name:32s} l2 / max rel err {rel(gw, gwd)}')
    except Exception as e:
        print(f'terms {terms} {name}: FAILED {e}')
