"""TEST INFRASTRUCTURE ONLY.  Stand-in for the two functions of POT ("Python Optimal Transport", environment.yml:38, no version pin) that the reference's
camera-adaptor regulariser calls (src/training/loss.py:195-197): `ot.dist` and `ot.emd2`.  POT is not installed in this image and cannot be fetched.

What POT documents for them, restated:
  * `ot.dist(x1, x2)` (default metric 'sqeuclidean'): M[i, j] = ||x1[i] - x2[j]||^2;
  * `ot.emd2(a, b, M)`: the optimal value  min_G <G, M>  over couplings G with marginals a, b (exact network-simplex solve).  With torch inputs the value is
    differentiable: the gradient with respect to M is the optimal coupling G itself (the plan is a constant of the backward pass).
For the uniform, equal-size marginals loss.py passes (a = b = 1/n), the vertices of the coupling polytope are permutation matrices / n, so the optimum is an
assignment problem, solved exactly here by scipy's Hungarian implementation -- an independent solver, NOT the product's sorted-matching shortcut
(3dgp_b200/training/loss.py::emd2_1d), which is the thing under test."""
import numpy as np
import torch
from scipy.optimize import linear_sum_assignment


def dist(x1, x2=None, metric='sqeuclidean'):
    assert metric == 'sqeuclidean'
    x2 = x1 if x2 is None else x2
    return (x1[:, None, :] - x2[None, :, :]).square().sum(dim=-1)


def emd2(a, b, M):
    n, m = M.shape
    assert n == m and torch.allclose(a, torch.full_like(a, 1.0 / n)) and torch.allclose(b, torch.full_like(b, 1.0 / n)), 'uniform equal-size marginals only'
    rows, cols = linear_sum_assignment(M.detach().cpu().double().numpy())
    plan = torch.zeros_like(M)
    plan[torch.as_tensor(rows), torch.as_tensor(cols)] = 1.0 / n
    return (plan * M).sum()
