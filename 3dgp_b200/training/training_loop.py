"""Per-iteration data path of the reference training loop (src/training/training_loop.py:296-366) around `step.Trainer`:

    real batch  <- BatchStream -> PrefetchLoader -> dataset.device_inputs          (:300-304: fetch, .to(device), /127.5 - 1, depth scaling)
    gen batch   <- z ~ N(0, I); conditioning labels of random training items (:305-308); cameras from the prior, or centred on the camera
                   angles of those items when the prior is `custom` (:309-313); camera-conditioning angles with a fraction `gpc_spoof_p`
                   rolled by one sample (:317-319)
    Trainer.step(real, gen)                                                        (:321-366: phases, all-reduce, Adam, G_ema)

Only this path is built.  The loop's control plane -- tick bookkeeping, snapshots, metrics, ADA, logging (:368-500) -- is out of scope.
One latent batch serves every phase of an iteration (the reference draws one per phase, :305 + :323); the phases are otherwise unchanged."""
import numpy as np
import torch

from ..dnnlib import EasyDict
from . import dataset as dataset_mod
from .rendering_utils import sample_camera_params


def sample_gen_batch(cfg, training_set, batch, device, z_dim, gpc_spoof_p=0.0, rng=np.random):
    """Generator-side inputs of one iteration (training_loop.py:305-319)."""
    z = torch.randn([batch, z_dim], device=device)
    pick = [rng.randint(len(training_set)) for _ in range(batch)]
    c = torch.from_numpy(np.stack([training_set.get_label(i) for i in pick])).to(torch.float32)
    c = (c.pin_memory() if device.type == 'cuda' else c).to(device, non_blocking=True)
    origin = None
    if cfg.camera.origin.angles.dist == 'custom':
        a = torch.from_numpy(np.stack([training_set.get_camera_angles(i) for i in pick])).to(torch.float32)
        origin = (a.pin_memory() if device.type == 'cuda' else a).to(device, non_blocking=True)
    cam = sample_camera_params(cfg.camera, batch, device, origin_angles=origin)
    cond = cam.angles.clone()
    if gpc_spoof_p > 0:
        spoof = (torch.rand(batch) < gpc_spoof_p).to(device)
        cond[spoof] = cond[spoof].roll(shifts=1, dims=0)
    return EasyDict(z=z, c=c, camera_params=cam, camera_angles_cond=cond)


def training_iterations(trainer, training_set, device, num_iters, batch=None, seed=0, workers=8, prefetch=2, gpc_spoof_p=None):
    """Generator over `num_iters` optimisation steps on `training_set` (a dataset.ImageFolderDataset); yields each step's stats dict.
    batch: images per rank (default cfg.training.batch_size // world_size, train.py:169).  Data are sharded by (rank, world_size) with the reference
    sampler's stream; rank r seeds it like training_loop.py:161."""
    from .inference import PrefetchLoader
    device = torch.device(device)
    cfg = trainer.cfg
    batch = batch or max(cfg.training.batch_size // trainer.world_size, 1)
    stream = dataset_mod.BatchStream(training_set, batch, rank=trainer.rank, num_replicas=trainer.world_size, seed=seed, workers=workers, depth=prefetch + 2)
    loader = PrefetchLoader(stream, device, depth=prefetch) if device.type == 'cuda' else stream
    z_dim = trainer.G.z_dim
    spoof = gpc_spoof_p if gpc_spoof_p is not None else float(getattr(trainer.loss, 'gpc_spoof_p', 0.0) or 0.0)
    try:
        for _ in range(num_iters):
            real = dataset_mod.device_inputs(next(loader))
            gen = sample_gen_batch(cfg, training_set, batch, device, z_dim, spoof)
            yield trainer.step(real, gen)
    finally:
        stream.close()
