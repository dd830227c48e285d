"""One data-parallel optimisation step of the reference training loop (src/training/training_loop.py:291-366):
phases Gmain / Dmain (+ lazy Dreg), gradient accumulation, ONE flattened all-reduce per phase (the only collective on
the hot path, :335-344) with the `/world + nan_to_num(+-1e5)` epilogue fused into one kernel, Adam, G_ema lerp.

One process per GPU; `torch.distributed` (NCCL over NVLink on the B200 box, gloo in CPU tests) carries the all-reduce.
"""
import copy

import torch
import torch.distributed as dist

from .. import _lib
from ..torch_utils.ops import tc as _tc
from ..dnnlib import EasyDict


def allreduce_gradients(params, world_size, group=None):
    """flat = cat(grads); all_reduce(SUM); flat = nan_to_num(flat / world, 0, 1e5, -1e5); scatter back (training_loop.py:335-344).
    On CUDA the divide + nan_to_num runs as one pass (gp3d_grad_epilogue) instead of two torch ops."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return 0
    flat = torch.cat([p.grad.flatten() for p in params])
    if world_size > 1:
        dist.all_reduce(flat, group=group)
    if flat.is_cuda:
        with torch.cuda.device(flat.device):
            rc = _lib.lib().gp3d_grad_epilogue(flat.data_ptr(), flat.numel(), 1.0 / world_size, 1e5, -1e5, _lib.stream_ptr())
        _lib.check(rc, 'grad_epilogue')
    else:   # host-side logic under test with gloo (no GPU): same arithmetic, torch ops
        flat = torch.nan_to_num(flat / world_size, nan=0, posinf=1e5, neginf=-1e5)
    for p, g in zip(params, flat.split([p.numel() for p in params])):
        p.grad = g.reshape(p.shape)
    return flat.numel()


class GradBuckets:
    """Bucketed all-reduce of a flat gradient buffer, overlapped with the backward pass that fills it (training_loop.py:335-344 issues ONE blocking
    all-reduce after backward; here the transfer hides behind the remaining backward kernels).

    The buffer is cut into contiguous buckets of whole parameters (built from its END, so that a bucket holds layers that finish together).  `arm()` is
    called right before the phase's FINAL backward; every parameter's post-accumulate hook then reports `ready(i)`, and a bucket whose parameters are all
    ready is all-reduced asynchronously on the process group's own stream (NCCL: ordered after the work already enqueued on the compute stream).
    Every rank must issue the same SEQUENCE of collectives, so buckets are launched strictly in a fixed `order`: by default their index order; in steady
    state the order in which they COMPLETED in the previous final backward of the same phase (`completion_order()`, a function of the static graph and
    therefore identical on every rank).  Storage order alone is not enough: G registers its mapping network after the synthesis network, so the tail of
    G's buffer -- bucket 0 -- holds the parameters whose gradients arrive LAST, and an index-ordered launch would hold every other bucket back until the
    backward is over.  `finish()` reduces what never completed (parameters without a gradient in this phase) and waits for all transfers."""

    def __init__(self, flat, offsets, numels, bucket_elems=16 << 20):
        self.flat = flat
        self.bounds, self.members = [], []          # bucket -> (start, end), [parameter indices]
        end, cur = flat.numel(), []
        for i in range(len(offsets) - 1, -1, -1):
            cur.append(i)
            if end - offsets[i] >= bucket_elems or i == 0:
                self.bounds.append((offsets[i], end)); self.members.append(cur)
                end, cur = offsets[i], []
        self.bucket_of = {i: b for b, m in enumerate(self.members) for i in m}
        self.armed = False
        self.launched_async = 0                     # buckets whose transfer started during the backward (diagnostics / tests)
        self.fire_sequence = []                     # parameter indices in the order their gradients arrived in the last armed backward

    def arm(self, world_size, group=None, expected=None, order=None):
        """expected: indices of the parameters this backward will produce gradients for (None: all).  A phase's graph is static, so the caller
        passes the set that `fired` in the previous final backward of the same phase; a bucket then does not wait for parameters that never fire.
        order: launch sequence of the buckets (a permutation of their indices; None: index order) -- the previous pass's `completion_order()`."""
        self.armed, self.world, self.group = True, world_size, group
        self.pending = [set(m) if expected is None else set(m) & set(expected) for m in self.members]
        self.order = list(range(len(self.bounds))) if order is None else list(order)
        assert sorted(self.order) == list(range(len(self.bounds))), 'order must be a permutation of the bucket indices'
        self.next, self.handles, self.launched_async, self.fired, self.fire_sequence, self.on_wire = 0, [], 0, set(), [], set()
        self.launch_log = []                        # (bucket, gradients arrived so far) per launch: how early each transfer started (diagnostics / tests)
        self._launch_ready()

    def _launch(self, b):
        lo, hi = self.bounds[b]
        self.handles.append(dist.all_reduce(self.flat[lo:hi], group=self.group, async_op=True))
        self.on_wire.add(b)
        self.launch_log.append((b, len(self.fire_sequence)))

    def _launch_ready(self):
        while self.next < len(self.order) and not self.pending[self.order[self.next]]:
            self._launch(self.order[self.next])
            self.next += 1
            self.launched_async += 1

    def ready(self, i):
        if self.armed:
            if i not in self.fired:
                self.fire_sequence.append(i)
            self.fired.add(i)
            b = self.bucket_of[i]
            if b in self.on_wire:   # its bucket is already on the wire: the sum would miss this rank's contribution
                raise RuntimeError('GradBuckets: a parameter outside the expected set produced a gradient after its bucket was all-reduced '
                                   '(the phase graph changed between iterations); set Trainer.overlap_allreduce = False')
            self.pending[b].discard(i)
            self._launch_ready()

    def completion_order(self):
        """Bucket indices sorted by when their last gradient arrived in the armed backward that just ran (buckets nothing fired for: first, they are
        complete from the start).  Deterministic for a static graph, hence the same on every rank."""
        at = {i: k for k, i in enumerate(self.fire_sequence)}
        done_at = [max((at[i] for i in m if i in at), default=-1) for m in self.members]
        return sorted(range(len(self.members)), key=lambda b: (done_at[b], b))

    def finish(self):
        """After the backward: reduce the remaining buckets, wait for every transfer (the compute stream then sees the reduced buffer)."""
        while self.next < len(self.order):
            self._launch(self.order[self.next])
            self.next += 1
        for h in self.handles:
            h.wait()
        self.handles, self.armed = [], False


class FlatAdam:
    """torch.optim.Adam over ONE flat float32 storage per module (training_loop.py:190-205, 335-346, 357-364).

    Parameters, gradients, second moments (first moments only when beta1 != 0) and, for G, the EMA copy live in parallel
    flat buffers carved in 1024-element blocks; each `p.data` / `p.grad` / `p_ema.data` is a view into them.  A phase is then
      zero the flat gradient -> backward accumulates in place -> ONE all-reduce of the flat gradient ->
      ONE kernel: /world + nan_to_num + Adam + p_ema lerp   (csrc/optim.cu)
    instead of cat + all_reduce + 2 elementwise passes + split/copy-back + ~12 foreach kernels + lerp/copy per tensor.
    Parameters that received no gradient in a phase are left untouched and keep their own step count, as torch.optim does
    for `.grad is None`."""
    BLOCK = 1024

    def __init__(self, module, lr, betas=(0.9, 0.999), eps=1e-8, ema_module=None, **unused):
        assert not unused.get('weight_decay') and not unused.get('amsgrad'), 'only the options the reference configs use are built'
        self.params = [p for p in module.parameters() if p.numel() > 0]
        dev = self.params[0].device
        assert dev.type == 'cuda' and all(p.dtype == torch.float32 and p.device == dev for p in self.params)
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += -(-p.numel() // self.BLOCK) * self.BLOCK
        self.numel = off
        self.flat_p = torch.zeros(off, device=dev)
        self.flat_g = torch.zeros(off, device=dev)
        self.flat_v = torch.zeros(off, device=dev)
        self.flat_m = torch.zeros(off, device=dev) if self.betas[0] != 0 else None
        self.flat_ema = torch.zeros(off, device=dev) if ema_module is not None else None
        self.grad_views = []
        for p, o in zip(self.params, self.offsets):
            v = self.flat_p[o:o + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
            self.grad_views.append(self.flat_g[o:o + p.numel()].view(p.shape))
        self._ema_ids = set()
        if ema_module is not None:
            ema_params = [p for p in ema_module.parameters() if p.numel() > 0]
            self._ema_ids = {id(p) for p in ema_params}
            assert [p.shape for p in ema_params] == [p.shape for p in self.params]
            for p, o in zip(ema_params, self.offsets):
                v = self.flat_ema[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
        seg = torch.full([off // self.BLOCK], -1, dtype=torch.int32)
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            seg[o // self.BLOCK:(o + p.numel() + self.BLOCK - 1) // self.BLOCK] = i
        self.blk_seg = seg.to(dev)
        self.steps = [0] * len(self.params)
        self._param_ids = {id(p) for p in self.params}
        self.active = set()
        self.buckets = GradBuckets(self.flat_g, self.offsets, [p.numel() for p in self.params])

        def hook(i):
            def fn(_p):
                self.active.add(i)
                self.buckets.ready(i)
            return fn
        self._hooks = [p.register_post_accumulate_grad_hook(hook(i)) for i, p in enumerate(self.params) if p.requires_grad]

    def zero_grad(self):
        """Start of a phase: clear the flat gradient and (re-)attach every parameter's .grad view."""
        self.flat_g.zero_()
        self.active.clear()
        for p, g in zip(self.params, self.grad_views):
            p.grad = g

    def step(self, world_size=1, group=None, ema_beta=None):
        """All-reduce the flat gradient (bucket transfers already in flight when the phase armed `buckets`), then the fused epilogue + Adam (+ EMA) kernel."""
        if self.buckets.armed:
            self.buckets.finish()
        elif world_size > 1:
            dist.all_reduce(self.flat_g, group=group)
        b1, b2 = self.betas
        for i in self.active:
            self.steps[i] += 1
        scal = lambda t: (self.lr / (1 - b1 ** t), (1 - b2 ** t) ** 0.5)
        uniform = len(self.active) == len(self.params) and len(set(self.steps)) == 1
        desc = None
        if uniform:
            step_size, bc2s = scal(self.steps[0])
        else:
            step_size, bc2s = 0.0, 1.0
            rows = []
            for i, t in enumerate(self.steps):
                a = i in self.active
                ss, bc = scal(t) if a else (0.0, 1.0)
                rows.append([ss, bc, 1.0 if a else 0.0, 0.0])
            desc = torch.tensor(rows, dtype=torch.float32).to(self.flat_p.device)
        with torch.cuda.device(self.flat_p.device):
            rc = _lib.lib().gp3d_adam_ema_step(self.flat_p.data_ptr(), self.flat_g.data_ptr(), _lib.ptr(self.flat_m), self.flat_v.data_ptr(),
                                               _lib.ptr(self.flat_ema) if ema_beta is not None else None, self.numel,
                                               1.0 / world_size, 1e5, -1e5, b1, b2, 1 - b1, 1 - b2, self.eps, step_size, bc2s,
                                               float(ema_beta) if ema_beta is not None else 0.0,
                                               None if desc is None else self.blk_seg.data_ptr(), _lib.ptr(desc), _lib.stream_ptr())
        _lib.check(rc, 'adam_ema_step')
        # the kernel wrote the parameters (and the EMA copy) behind autograd's version counters: drop their cached bf16 conv operands
        _tc.invalidate_weight_cache(self._param_ids | self._ema_ids if ema_beta is not None else self._param_ids)
        self.active.clear()
        return self.numel

    def state_dict(self):
        """Per-parameter state in torch.optim.Adam's layout (step, exp_avg, exp_avg_sq), for checkpoints."""
        st = {}
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            sl = slice(o, o + p.numel())
            st[i] = dict(step=self.steps[i], exp_avg_sq=self.flat_v[sl].view(p.shape).clone(),
                         exp_avg=(self.flat_m[sl].view(p.shape).clone() if self.flat_m is not None else None))
        return dict(state=st, lr=self.lr, betas=self.betas, eps=self.eps)

    def load_state_dict(self, sd):
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            e = sd['state'][i]; sl = slice(o, o + p.numel())
            self.steps[i] = int(e['step'])
            self.flat_v[sl].copy_(e['exp_avg_sq'].flatten())
            if self.flat_m is not None:
                self.flat_m[sl].copy_(e['exp_avg'].flatten())


class Trainer:
    """Holds G, D, G_ema, the loss and both Adam optimisers with the reference's lazy-regularisation scaling
    (training_loop.py:190-205: lr *= mb_ratio, betas ** mb_ratio with mb_ratio = interval / (interval + 1))."""

    def __init__(self, G, D, loss, cfg, rank=0, world_size=1, D_reg_interval=16, ema_kimg=10.0, ema_rampup=0.05, batch_size=None, micro_batch=None):
        self.G, self.D, self.loss, self.cfg, self.rank, self.world_size = G, D, loss, cfg, rank, world_size
        if world_size > 1:   # training_loop.py:176-179: every rank starts from rank 0's parameters and buffers
            for module in (G, D):
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t.data, src=0)
        self.G_ema = copy.deepcopy(G).eval().requires_grad_(False)
        gk, dk = dict(cfg.model.generator.optim.kwargs), dict(cfg.model.discriminator.optim.kwargs)
        mb = D_reg_interval / (D_reg_interval + 1) if D_reg_interval else 1.0
        dk['lr'] = dk['lr'] * mb
        dk['betas'] = [b ** mb for b in dk['betas']]
        self.flat = next(G.parameters()).is_cuda      # CUDA: flat storage + fused kernel; CPU (gloo host-logic tests): torch ops
        if self.flat:
            self.G_opt = FlatAdam(G, ema_module=self.G_ema, **gk)
            self.D_opt = FlatAdam(D, **dk)
        else:
            self.G_opt = torch.optim.Adam(G.parameters(), **gk)
            self.D_opt = torch.optim.Adam(D.parameters(), **dk)
        self.D_reg_interval = D_reg_interval
        self.ema_kimg, self.ema_rampup, self.batch_size = ema_kimg, ema_rampup, batch_size
        self.micro_batch = micro_batch
        # Bucketed transfer overlapped with the final backward (GradBuckets).  OFF by default: validated on 2 ranks (gloo on CPU, NCCL on 2 x B200: same
        # step time as the blocking form) but a 4-rank NCCL run did not complete (profiles/r2_allreduce_overlap.txt) -- the blocking flat all-reduce
        # of training_loop.py:335-344 is the shipped schedule.
        self.overlap_allreduce = False
        self._fired = {}         # phase -> parameters that produced gradients in its last final backward (GradBuckets.arm `expected`)
        self._order = {}         # phase -> bucket launch sequence = completion order of that backward (GradBuckets.arm `order`)
        self.cur_nimg = 0
        self.it = 0

    def _phase(self, name, module, opt, real, gen, gain, render_opts=None, ema_beta=None):
        if self.flat:
            opt.zero_grad()
        else:
            opt.zero_grad(set_to_none=True)
        module.requires_grad_(True)
        stats = {}
        mbs = list(self._micro_batches(real, gen))
        for j, (r_mb, g_mb) in enumerate(mbs):                 # gradient accumulation, training_loop.py:329-330
            # the last backward of the phase streams its gradient buckets into the all-reduce while it is still running
            # (never in the first two iterations: first-use kernel loading / allocator growth must not interleave with collectives already spinning on peers)
            arm = (lambda: opt.buckets.arm(self.world_size, expected=self._fired.get(name), order=self._order.get(name))) if (self.flat and self.world_size > 1 and self.overlap_allreduce
                                                                                                     and self.it >= 2 and j == len(mbs) - 1) else None
            stats = self.loss.accumulate_gradients(phase=name, real_data=r_mb, gen_data=g_mb, gain=gain, cur_nimg=self.cur_nimg, render_opts=render_opts,
                                                   final_backward=arm)
        module.requires_grad_(False)
        if self.flat:
            armed = opt.buckets.armed
            opt.step(self.world_size, ema_beta=ema_beta)
            if armed:
                self._fired[name] = set(opt.buckets.fired)
                self._order[name] = opt.buckets.completion_order()
        else:
            allreduce_gradients([p for p in module.parameters() if p.numel() > 0], self.world_size)
            opt.step()
        return stats

    def _micro_batches(self, real, gen):
        n = len(gen.z)
        mb = self.micro_batch or n
        if mb >= n:
            yield real, gen
            return
        sl = lambda d, a, b: EasyDict(**{k: (v[a:b] if hasattr(v, '__getitem__') else v) for k, v in d.items()})
        for a in range(0, n, mb):
            yield sl(real, a, a + mb), sl(gen, a, a + mb)

    def step(self, real, gen, render_opts=None):
        """real/gen: EasyDicts of this rank's micro-batch (see loss.accumulate_gradients).  Returns scalar stats."""
        stats = {}
        # G_ema coefficient (training_loop.py:357-362); G does not change after Gmain, so the lerp rides in G's optimiser kernel
        bs = self.batch_size or (len(gen.z) * self.world_size)
        ema_nimg = self.ema_kimg * 1000
        if self.ema_rampup is not None:
            ema_nimg = min(ema_nimg, self.cur_nimg * self.ema_rampup)
        beta = 0.5 ** (bs / max(ema_nimg, 1e-8))
        self.D.requires_grad_(False)
        stats.update(self._phase('Gmain', self.G, self.G_opt, real, gen, 1, render_opts, ema_beta=beta))
        self.G.requires_grad_(False)
        stats.update(self._phase('Dmain', self.D, self.D_opt, real, gen, 1, render_opts))
        if self.D_reg_interval and self.it % self.D_reg_interval == 0:
            stats.update(self._phase('Dreg', self.D, self.D_opt, real, gen, self.D_reg_interval, render_opts))
        # G_ema (training_loop.py:357-366)
        with torch.no_grad():
            if not self.flat:
                for pe, p in zip(self.G_ema.parameters(), self.G.parameters()):
                    pe.copy_(p.lerp(pe, beta))
            for be, b in zip(self.G_ema.buffers(), self.G.buffers()):
                be.copy_(b)
        self.cur_nimg += bs
        self.it += 1
        return stats
