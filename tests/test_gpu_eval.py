"""Eval-side pipeline on the GPU: uint8 conversion kernel, the FID2k generator call pattern (metric_utils.py:303-319) and the host -> device prefetcher."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import cases

pytestmark = pytest.mark.gpu


def _same_images(a, b):
    """uint8 images of two runs of the SMALL networks: their 32-channel convolutions run on cuDNN (below the tensor-core kernels' channel granularity),
    whose fp32 algorithms are not run-to-run deterministic (measured: 25 of 49 152 float pixels differ by <= 1e-3 between two eager calls,
    tools/debug_graph.py) -- a handful of pixels may sit on a rounding boundary.  Everything this repository launches is deterministic."""
    d = (a.int() - b.int()).abs()
    return int(d.max()) <= 1 and float((d > 0).float().mean()) < 2e-3


def _inf():
    return importlib.import_module('3dgp_b200.training.inference')


@pytest.mark.parametrize('cl', [False, True])
@pytest.mark.parametrize('shape', [(2, 4, 16, 16), (3, 3, 8, 20), (1, 4, 64, 64)])
def test_to_uint8_equals_the_reference_expression(shape, cl):
    """(img[:, :3] * 127.5 + 128).clamp(0, 255).to(torch.uint8) (metric_utils.py:313), bit for bit, NCHW and channels-last inputs."""
    torch.manual_seed(shape[2])
    x = torch.randn(shape, device='cuda') * 1.5
    x[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, 0.99999, -5.0], device='cuda')
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    y = _inf().to_uint8(x)
    ref = (x[:, :3] * 127.5 + 128).clamp(0, 255).to(torch.uint8)
    assert y.is_contiguous() and torch.equal(y, ref)


def test_generate_uint8_runs_a_snapshot_generator(golden):
    """Snapshot pickle (reference persistence format) -> this package's G_ema -> the metrics loop's generator call -> uint8 images on the fused path."""
    lg = importlib.import_module('3dgp_b200.legacy')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    Ge = lg.load_network_pkl(os.path.join(ROOT, 'tests', 'golden', 'snapshot_small.pkl.gz'), device='cuda', names=('G_ema',))['G_ema']
    kw = cases.net_kwargs('small')
    t = {k: torch.from_numpy(v).cuda() for k, v in cases.net_inputs(kw).items()}
    cam = dn.TensorGroup(angles=t['angles'], fov=t['fov'], radius=t['radius'], look_at=t['look_at'])
    Ge.synthesis.renderer.launch_counter = 0          # the stratification jitter is a fresh Philox stream per launch: replay the same one twice
    img = _inf().generate_uint8(Ge, t['z'], t['c'], cam, noise_mode='const')
    B = t['z'].shape[0]
    assert img.dtype == torch.uint8 and tuple(img.shape) == (B, 3, kw['img_resolution'], kw['img_resolution'])
    Ge.synthesis.renderer.launch_counter = 0
    with torch.no_grad():
        ref = Ge(z=t['z'], c=t['c'], camera_params=cam, camera_angles_cond=cam.angles, noise_mode='const')
    ref = ref if torch.is_tensor(ref) else ref.img
    assert _same_images(img, (ref[:, :3] * 127.5 + 128).clamp(0, 255).to(torch.uint8))


def test_prefetch_loader_delivers_every_batch_in_order():
    inf = _inf()
    batches = [dict(img=torch.full((4, 3, 8, 8), float(i)), c=torch.arange(4) + 10 * i) for i in range(5)]
    got = list(inf.PrefetchLoader(iter(batches), 'cuda', depth=2))
    assert len(got) == 5
    for i, b in enumerate(got):
        assert b['img'].is_cuda and torch.equal(b['img'].cpu(), batches[i]['img']) and torch.equal(b['c'].cpu(), batches[i]['c'])


def test_training_set_stream_reaches_the_device_unchanged():
    """training/dataset.py::BatchStream (worker threads decoding into pinned batch buffers, the reference sampler's order) -> PrefetchLoader (H2D on a side
    stream) -> device_inputs (training_loop.py:300-304 on the device): every delivered batch equals the items the dataset returns on the host, although
    the ring's pinned buffers are refilled while earlier copies are in flight."""
    inf = _inf()
    dsmod = importlib.import_module('3dgp_b200.training.dataset')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    cfg = dn.EasyDict.init_recursively(dict(c_dim=3, use_embeddings=False, mirror=True, camera=dict(
        fov=dict(dist='uniform', min=10.0, max=45.0), origin=dict(radius=dict(dist='normal', mean=1.0, std=0.0),
        angles=dict(dist='uniform', yaw=dict(min=-1.57, max=1.57), pitch=dict(min=0.785398163, max=2.35619449))))))
    ds = dsmod.ImageFolderDataset(path=os.path.join(ROOT, 'tests', 'golden', 'tiny_dataset.zip'), cfg=cfg)
    stream = dsmod.BatchStream(ds, batch=6, seed=2, workers=4, depth=4)
    assert stream.ring[0]['image'].is_pinned()
    order = dsmod.infinite_order(len(ds), 0, 1, True, 2)
    loader = inf.PrefetchLoader(stream, 'cuda', depth=2)
    for _ in range(9):
        b = next(loader)
        idx = [next(order) for _ in range(6)]
        x = dsmod.device_inputs(b)
        assert x.img.is_cuda and x.img.dtype == torch.float32
        want = np.stack([ds[i]['image'] for i in idx])
        assert torch.equal(b['image'].cpu(), torch.from_numpy(want))
        # ATen divides by a scalar as a multiplication by its reciprocal on CUDA: one ulp from the host's true division (measured, GPU pass AG)
        assert torch.allclose(x.img.cpu(), torch.from_numpy(want).to(torch.float32) / 127.5 - 1.0, rtol=0, atol=5e-7)
        assert torch.equal(b['label'].cpu(), torch.from_numpy(np.stack([ds[i]['label'] for i in idx])))
        assert torch.equal(b['camera_angles'].cpu(), torch.from_numpy(np.stack([ds[i]['camera_angles'] for i in idx])))
    stream.close()


def test_cuda_graph_replay_of_the_generator_equals_the_eager_call():
    """training/inference.py::GraphedGenerator: the captured generator call replayed on new latents / cameras returns the eager result (same Philox launch offset;
    bit for bit up to cuDNN's own run-to-run differences, see _same_images), for several batches in a row."""
    lg = importlib.import_module('3dgp_b200.legacy')
    dn = importlib.import_module('3dgp_b200.dnnlib')
    inf = _inf()
    Ge = lg.load_network_pkl(os.path.join(ROOT, 'tests', 'golden', 'snapshot_small.pkl.gz'), device='cuda', names=('G_ema',))['G_ema']
    kw = cases.net_kwargs('small')
    t = {k: torch.from_numpy(v).cuda() for k, v in cases.net_inputs(kw).items()}
    B = t['z'].shape[0]
    gg = inf.GraphedGenerator(Ge, B, noise_mode='const')
    captured_offset = Ge.synthesis.renderer.launch_counter            # the graph baked this launch's Philox offset in
    for rep in range(3):
        z = t['z'] + 0.1 * rep
        cam = dn.TensorGroup(angles=t['angles'] + 0.01 * rep, fov=t['fov'], radius=t['radius'], look_at=t['look_at'])
        out = gg(z, t['c'], cam).clone()
        Ge.synthesis.renderer.launch_counter = captured_offset - 1
        ref = inf.generate_uint8(Ge, z, t['c'], cam, noise_mode='const')
        assert _same_images(out, ref), rep
