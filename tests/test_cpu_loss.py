"""Host-side loss arithmetic that needs no GPU: the earth-mover regulariser of the camera adaptor (reference src/training/loss.py:182-216).

The reference computes it with POT (`ot.dist` + `ot.emd2`), a dependency that is neither vendored in the reference nor installed here.  For uniform
weights and equal sample counts the transport problem `emd2` solves exactly is an assignment problem, so an independent exact solver of the SAME
cost matrix (scipy's Hungarian-type `linear_sum_assignment`) pins the value, and the gradient POT returns (the optimal plan applied to d(cost))
follows from that assignment."""
import importlib

import numpy as np
import pytest
import torch

loss_mod = importlib.import_module('3dgp_b200.training.loss')


def _exact_emd2(a, b):
    """emd2(1/n, 1/n, sqeuclidean dist(a[:, [i]], b[:, [i]])) per column through an exact assignment; returns (values [k], d values / d a [n, k])."""
    from scipy.optimize import linear_sum_assignment
    n, k = a.shape
    val = np.zeros(k); grad = np.zeros_like(a)
    for i in range(k):
        M = (a[:, i][:, None] - b[:, i][None, :]) ** 2              # ot.dist default metric: squared Euclidean
        r, c = linear_sum_assignment(M)
        val[i] = M[r, c].sum() / n                                   # plan = permutation / n
        grad[r, i] = 2.0 * (a[r, i] - b[c, i]) / n
    return val, grad


@pytest.mark.parametrize('n,seed', [(64, 0), (33, 1), (256, 2)])
def test_emd_restatement_equals_the_exact_assignment_solution(n, seed):
    rs = np.random.RandomState(seed)
    a = rs.randn(n, 8) * rs.uniform(0.1, 3.0, size=[1, 8]) + rs.randn(1, 8)
    b = rs.randn(n, 8)
    a[:, 5] = a[:, 5].round(1)                                       # ties inside one sample: the optimum is degenerate, its VALUE is not
    at = torch.tensor(a, dtype=torch.float64, requires_grad=True)
    got = loss_mod.emd2_1d(at, torch.tensor(b, dtype=torch.float64))
    assert tuple(got.shape) == (1, 8)
    val, grad = _exact_emd2(a, b)
    np.testing.assert_allclose(got.detach().numpy()[0], val, rtol=1e-12, atol=1e-14)
    got.sum().backward()
    cols = [i for i in range(8) if i != 5]                           # unique optimal plan (continuous data): the gradient is determined
    np.testing.assert_allclose(at.grad.numpy()[:, cols], grad[:, cols], rtol=1e-10, atol=1e-13)


def test_emd_of_a_sample_with_itself_is_zero_and_shift_is_quadratic():
    x = torch.randn(50, 8, dtype=torch.float64)
    assert float(loss_mod.emd2_1d(x, x[torch.randperm(50)]).abs().max()) == 0.0
    np.testing.assert_allclose(loss_mod.emd2_1d(x + 0.5, x).numpy(), np.full([1, 8], 0.25), rtol=1e-12)


def test_loss_takes_the_reference_constructor_keywords_and_phase_names():
    """training_loop.py:186 builds the loss as `construct_class_by_name(device=, G=, D=, augment_pipe=, cfg=, r1_gamma=)` and, without lazy G regularisation,
    runs G's phase under the name 'Gall' (:195): the stand-alone loss accepts both; switches that are off on the 3dgp path refuse loudly."""
    import importlib
    import pytest
    import torch
    cfgm = importlib.import_module('3dgp_b200.config')
    lossm = importlib.import_module('3dgp_b200.training.loss')
    cfg = cfgm.make_config(cmax=32, cbase=512, tri_res=16, patch_res=8, img_resolution=16, c_dim=4, depth_hid=8, hid_dim=16, w_dim=32, z_dim=32, embedding_dim=8, num_ray_steps=4, use_depth=False)      # no depth channel: run_D then needs no FIR launch
    G, D = torch.nn.Module(), torch.nn.Module()
    calls = []

    class Pipe:
        def __call__(self, img, num_color_channels):
            calls.append(num_color_channels); return img * 2
    G.img_channels = 3
    D.forward = lambda img, c, update_emas=False, **kw: (img.sum([1, 2, 3]), None)
    L = lossm.StyleGAN2Loss(device='cpu', G=G, D=D, augment_pipe=Pipe(), cfg=cfg, r1_gamma=0.5, style_mixing_prob=0, pl_batch_shrink=2, pl_decay=0.01)
    img = torch.ones(2, 3, 8, 8)
    logits, _ = L.run_D(img, None, blur_sigma=0)
    assert calls == [3] and torch.equal(logits, torch.full([2], 2.0 * 3 * 64))            # the pipe saw the image, D saw the pipe's output
    with pytest.raises(NotImplementedError):
        lossm.StyleGAN2Loss(cfg, 'cpu', G, D, style_mixing_prob=0.9)
    seen = []
    L.run_G = lambda *a, **k: (_ for _ in ()).throw(RuntimeError(seen.append('Gmain') or 'reached'))
    with pytest.raises(RuntimeError, match='reached'):
        L.accumulate_gradients(phase='Gall', real_data=cfgm.EasyDict(img=img, depth=img[:, :1]), gen_data=cfgm.EasyDict(z=None, c=None, camera_params=None), gain=1, cur_nimg=0)
    assert seen == ['Gmain']
