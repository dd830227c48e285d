"""Eval-side generator path (the FID2k call pattern, reference src/metrics/metric_utils.py:288-319, and the snapshot image grid of
training_loop.py:23-49): G in eval mode, full-frame render at img_resolution, images converted to uint8 on the device by ONE kernel
(csrc/misc.cu: gp3d_to_uint8) instead of four elementwise passes, optionally copied to pinned host memory asynchronously."""
import torch

from .. import _lib


def to_uint8(img, channels=3, scale=127.5, shift=128.0):
    """(img[:, :channels] * scale + shift).clamp(0, 255).to(torch.uint8) -- metric_utils.py:313 -- for a float32 [N, C, H, W] tensor of any strides
    (NCHW or channels-last), W % 4 == 0.  Returns NCHW-contiguous uint8."""
    L = _lib.lib()
    _lib.require_cuda(img, 'img')
    if img.dtype != torch.float32 or img.dim() != 4:
        raise RuntimeError('to_uint8: img must be a float32 [N, C, H, W] tensor')
    N, C, H, W = img.shape
    out = torch.empty([N, channels, H, W], dtype=torch.uint8, device=img.device)
    with torch.cuda.device(img.device):
        rc = L.gp3d_to_uint8(img.data_ptr(), out.data_ptr(), N, C, channels, H, W, img.stride(0), img.stride(1), img.stride(2), img.stride(3),
                             float(scale), float(shift), _lib.stream_ptr())
    _lib.check(rc, 'to_uint8')
    return out


@torch.no_grad()
def generate_uint8(G, z, c, camera_params, **G_kwargs):
    """One generator batch of the metrics loop (metric_utils.py:306-313): optional camera adaptor, G(z, c, camera), uint8 RGB.
    G must be in eval mode (G_ema); G_kwargs as opts.G_kwargs (e.g. noise_mode='const')."""
    cad = getattr(G.synthesis, 'camera_adaptor', None)
    if cad is not None and G.cfg.camera_adaptor.enabled:
        camera_params = cad(camera_params, z, c)
    img = G(z=z, c=c, camera_params=camera_params, camera_angles_cond=camera_params.angles, **G_kwargs)
    if not torch.is_tensor(img):
        img = img.img
    return to_uint8(img, channels=min(3, img.shape[1]))


class PrefetchLoader:
    """Host -> device prefetch for an iterator of dict-of-CPU-tensor batches (the role of the reference's DataLoader(pin_memory=True, prefetch_factor=2)
    + `.to(device)` at the top of every iteration, training_loop.py:160-166, 300-312): batch i + 1 is staged in pinned memory and copied on a side stream
    while batch i is being consumed; `next()` makes the compute stream wait on the copy's event only."""

    def __init__(self, it, device, depth=2):
        self.it, self.device, self.depth = iter(it), torch.device(device), max(int(depth), 1)
        self.stream = torch.cuda.Stream(self.device)
        self.queue = []
        for _ in range(self.depth):
            self._stage()

    def _stage(self):
        try:
            host = next(self.it)
        except StopIteration:
            return
        pinned = {k: (v if v.is_pinned() else v.pin_memory()) for k, v in host.items()}
        with torch.cuda.stream(self.stream):
            dev = {k: v.to(self.device, non_blocking=True) for k, v in pinned.items()}
            ev = torch.cuda.Event(); ev.record(self.stream)
        self.queue.append((dev, ev, pinned))      # the pinned source must outlive the asynchronous copy

    def __iter__(self):
        return self

    def __next__(self):
        if not self.queue:
            raise StopIteration
        dev, ev, _pinned = self.queue.pop(0)
        # host wait first: once a batch is handed out its pinned source is free again -- a producer that recycles a ring of pinned buffers
        # (training/dataset.py::BatchStream, ring length >= depth + 1) may refill it.  The copy was issued `depth` iterations ago: this does not stall.
        ev.synchronize()
        torch.cuda.current_stream(self.device).wait_event(ev)
        for v in dev.values():
            v.record_stream(torch.cuda.current_stream(self.device))
        self._stage()
        return dev


class GraphedGenerator:
    """CUDA-graph replay of `generate_uint8` for a fixed batch size: the metrics loop of the reference generates images in batches of
    `batch_gen = 4` (metric_utils.py:289), where one G forward is ~2 000 kernel launches of a few microseconds each -- launch-bound on the host.
    The whole call (mapping network, camera adaptor, tri-plane decoder, ray generation + march, depth adaptor, uint8 conversion) is captured once
    on static input buffers and replayed; inputs are copied into the buffers, the result is the static uint8 output (clone it to keep it).
    The stratification jitter is a Philox stream keyed by a host-side launch counter, which a graph bakes in: every replay draws the SAME jitter
    pattern (latents and cameras still differ per batch); pass `fresh_jitter=True` to `__call__` to fall back to the eager path for a call."""

    def __init__(self, G, batch, device=None, warmup=2, **G_kwargs):
        from ..dnnlib import TensorGroup
        self.G, self.G_kwargs = G, G_kwargs
        dev = torch.device(device) if device is not None else next(G.parameters()).device
        f32 = dict(dtype=torch.float32, device=dev)
        self.z = torch.zeros([batch, G.z_dim], **f32)
        self.c = torch.zeros([batch, G.c_dim], **f32)
        self.cam = TensorGroup(angles=torch.zeros([batch, 3], **f32), fov=torch.full([batch], 20.0, **f32), radius=torch.ones([batch], **f32),
                               look_at=torch.zeros([batch, 3], **f32))
        self.cam.angles[:, 1] = 1.5707963
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                    # warm-up outside the capture: lazy initialisation, weight-operand caches, allocator pools
            for _ in range(max(int(warmup), 1)):
                generate_uint8(G, self.z, self.c, self.cam, **G_kwargs)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = generate_uint8(G, self.z, self.c, self.cam, **G_kwargs)

    @torch.no_grad()
    def __call__(self, z, c, camera_params, fresh_jitter=False):
        if fresh_jitter:
            return generate_uint8(self.G, z, c, camera_params, **self.G_kwargs)
        self.z.copy_(z, non_blocking=True); self.c.copy_(c, non_blocking=True)
        for k in ('angles', 'fov', 'radius', 'look_at'):
            self.cam[k].copy_(camera_params[k], non_blocking=True)
        self.graph.replay()
        return self.out
