"""Training-set reader and host -> device batch stream ("synthetic / zip dataset loader with H2D prefetch").

Reads the datasets the reference's `dataset_tool.py` writes -- a directory or a zip of images with an optional `dataset.json`
({"labels": [[fname, int | [float...]]...], "camera_angles": [[fname, [yaw, pitch, roll]]...]}) and optional `<name>_depth.png` maps -- and presents
the surface `train.py` / `training_loop.py` use (reference src/training/dataset.py:29-362): class name `ImageFolderDataset`, its constructor
arguments, the item dictionary {'image', 'label', 'camera_angles', 'depth', 'embedding'} and the properties the loop reads (`image_shape`,
`resolution`, `label_dim`, `has_labels`, `mean_camera_params`, `get_label`, `get_camera_angles`, ...).  Integer / byte work: outputs are bit-identical
to the reference's on the same files (tests/test_cpu_dataset.py, golden written by the unmodified reference).

What is different underneath (this is the B200 side of the data path, training_loop.py:160-166 + :300-312):
  * `infinite_order` restates `InfiniteSampler` (src/torch_utils/misc.py:112-143) as a plain generator, bit-for-bit the same index sequence;
  * `BatchStream` replaces `DataLoader(pin_memory=True, prefetch_factor=2)` + per-iteration `.to(device)`: worker THREADS (PNG / zip decoding releases
    the GIL) decode straight into the rows of preallocated pinned batch buffers -- no per-sample tensors, no collate copy, no pickling between
    processes -- and a ring of such buffers feeds `training/inference.py::PrefetchLoader`, which issues the asynchronous H2D copy of batch i + 1 on a
    side stream while batch i trains.  180 GB of HBM leave room to normalise on the device: `device_inputs` applies `/127.5 - 1` (and the depth
    scaling of :303) to the uint8 / int32 batch after the copy, so the PCIe transfer carries 1 byte per sample instead of 4.
Embeddings (`cfg.use_embeddings`, a memmap of pre-extracted features used by the knowledge-distillation term) are read the reference's way when
configured; depth maps are decoded with PIL (the reference uses pyspng, which is not installed here; the items are pinned against the reference reader run
with a specification-level PNG decoder in pyspng's place, tests/golden/dataset_depth_golden.npz)."""
import json
import os
import threading
import zipfile
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from .rendering_utils import get_mean_angles_values, get_mean_sampling_value

_IMAGE_EXT = None


def _image_extensions():
    global _IMAGE_EXT
    if _IMAGE_EXT is None:
        import PIL.Image
        PIL.Image.init()
        _IMAGE_EXT = set(PIL.Image.EXTENSION)
    return _IMAGE_EXT


def _ext(name):
    return os.path.splitext(name)[1].lower()


def _strip_root(fname, root):
    """Keys of dataset.json are relative to the archive's root folder (dataset.py:365-375)."""
    for prefix in (root + '/', '/' + root + '/'):
        if fname.startswith(prefix):
            return fname[len(prefix):]
    return '' if fname in (root, '/' + root) else fname


class _Archive:
    """File listing + byte access for a directory or a zip.  Zip handles are per thread (`zipfile.ZipFile` serialises concurrent reads on one handle) and per process (never shared across fork)."""

    def __init__(self, path):
        self.path = path
        self._local = threading.local()
        if os.path.isdir(path):
            self.kind = 'dir'
            self.names = {os.path.relpath(os.path.join(r, f), start=path) for r, _d, fs in os.walk(path) for f in fs}
        elif _ext(path) == '.zip':
            self.kind = 'zip'
            with zipfile.ZipFile(path) as z:
                self.names = set(z.namelist())
        else:
            raise IOError(f'Path must point to a directory or zip, but got {path}.')

    def read(self, name):
        if self.kind == 'dir':
            with open(os.path.join(self.path, name), 'rb') as f:
                return f.read()
        z = getattr(self._local, 'zip', None)
        if z is None or getattr(self._local, 'pid', None) != os.getpid():
            # first use in this thread -- or in this PROCESS: a handle inherited through fork() (DataLoader workers) shares its file offset with the
            # parent and with every sibling, and concurrent seeks corrupt each other's reads
            z = self._local.zip = zipfile.ZipFile(self.path)
            self._local.pid = os.getpid()
        return z.read(name)

    def close(self):
        z = getattr(self._local, 'zip', None)
        if z is not None:
            z.close()
            self._local.zip = None

    def __getstate__(self):                 # picklable for multi-process consumers: handles are re-opened lazily
        return dict(path=self.path, kind=self.kind, names=self.names)

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._local = threading.local()


def _decode_image(data):
    import io
    import PIL.Image
    a = np.array(PIL.Image.open(io.BytesIO(data)))
    if a.ndim == 2:
        a = a[:, :, None]
    return a.transpose(2, 0, 1)             # CHW


def _decode_depth(data):
    """First channel of an 8- or 16-bit depth PNG as int32 [1, h, w]; 8-bit maps are scaled to the 16-bit range (dataset.py:310-323)."""
    import io
    import PIL.Image
    im = PIL.Image.open(io.BytesIO(data))
    a = np.array(im)
    if a.ndim == 3:
        a = a[:, :, 0]
    if a.dtype == np.uint8:
        a = a.astype(np.uint16) * 256
    elif a.dtype not in (np.uint16, np.int32):      # PIL hands 16-bit greyscale out as uint16 ('I;16') or int32 ('I')
        raise IOError(f'Unsupported depth dtype {a.dtype}')
    return a.astype(np.int32)[None]


class ImageFolderDataset(torch.utils.data.Dataset):
    """Same constructor and item format as the reference class of this name (dataset.py:242-270, 29-61)."""

    def __init__(self, path, resolution=None, max_size=None, use_depth=False, random_seed=0, cfg=None, name=None, **_ignored):
        self.cfg = cfg if cfg is not None else {}
        self._path = path
        self._arc = _Archive(path)
        ext = _image_extensions()
        self._image_fnames = sorted(n for n in self._arc.names if _ext(n) in ext and not n.endswith('_depth.png'))
        if not self._image_fnames:
            raise IOError('No image files found in the specified path')
        self._name = name or os.path.splitext(os.path.basename(path.rstrip('/')))[0]
        first = self._load_raw_image(0)
        self._raw_shape = [len(self._image_fnames)] + list(first.shape)
        if resolution is not None and (self._raw_shape[2] != resolution or self._raw_shape[3] != resolution):
            raise IOError('Image files do not match the specified resolution')
        g = (lambda k, d=None: self.cfg.get(k, d)) if hasattr(self.cfg, 'get') else (lambda k, d=None: d)
        self._use_labels = (g('c_dim', 0) or 0) > 0
        self._use_embeddings = bool(g('use_embeddings', False))
        self._use_depth = bool(use_depth)
        self._labels = self._angles = self._emb = self._mean_cam = self._label_shape = None
        # subset (before the mirror, dataset.py:52-55) and horizontal mirror (:58-61)
        idx = np.arange(self._raw_shape[0], dtype=np.int64)
        if max_size is not None and idx.size > max_size:
            np.random.RandomState(random_seed).shuffle(idx)
            idx = np.sort(idx[:max_size])
        flip = np.zeros(idx.size, dtype=np.uint8)
        if g('mirror', False):
            idx = np.tile(idx, 2)
            flip = np.concatenate([flip, np.ones_like(flip)])
        self._raw_idx, self._xflip = idx, flip

    # ---- raw access -----------------------------------------------------------------------------------------------------------
    def _load_raw_image(self, raw_idx):
        return _decode_image(self._arc.read(self._image_fnames[raw_idx]))

    def _load_raw_depth(self, raw_idx):
        fname = self._image_fnames[raw_idx]
        return _decode_depth(self._arc.read(fname[:-len(_ext(fname))] + '_depth.png'))

    def _field(self, key):
        metas = [n for n in self._arc.names if n.endswith('dataset.json')]
        if not metas:
            return None
        assert len(metas) == 1, 'There can be only a single dataset.json file'
        values = json.loads(self._arc.read(metas[0])).get(key)
        if values is None:
            return None
        table = dict(values)
        return np.array([table[_strip_root(n, self._name).replace('\\', '/')] for n in self._image_fnames])

    def _raw_labels(self):
        if self._labels is None:
            lab = self._field('labels') if self._use_labels else None
            if lab is None:
                assert not self._use_labels, "We planned to use labels, but couldn't load them from dataset.json"
                lab = np.zeros([self._raw_shape[0], 0], dtype=np.float32)
            else:
                lab = lab.astype({1: np.int64, 2: np.float32}[lab.ndim])
            assert lab.shape[0] == self._raw_shape[0]
            if lab.dtype == np.int64:
                assert lab.ndim == 1 and np.all(lab >= 0)
            self._labels = lab
        return self._labels

    def _raw_camera_angles(self):
        if self._angles is None:
            a = self._field('camera_angles')
            self._angles = np.zeros([self._raw_shape[0], 3], dtype=np.float32) if a is None else a.astype(np.float32)
            assert self._angles.shape[0] == self._raw_shape[0]
        return self._angles

    def _raw_embeddings(self):
        if self._emb is None:
            if self._use_embeddings:
                with open(self.cfg.embeddings_desc_path) as f:
                    desc = json.load(f)
                emb = np.memmap(self.cfg.embeddings_path, dtype='float32', mode='r', shape=tuple(desc['shape']))
                rows = np.array([desc['filepath_to_idx'][_strip_root(n, self._name).replace('\\', '/')] for n in self._image_fnames]).astype(np.int32)
            else:
                emb, rows = np.zeros([self._raw_shape[0], 0], dtype=np.float32), np.arange(self._raw_shape[0])
            self._emb = (rows, emb)
        return self._emb

    # ---- items ----------------------------------------------------------------------------------------------------------------
    def __len__(self):
        return self._raw_idx.size

    def image_into(self, idx, out):
        """Decodes item `idx` (mirror applied) into `out` [C, H, W] uint8 -- a row of a pinned batch buffer."""
        img = self._load_raw_image(self._raw_idx[idx])
        if list(img.shape) != self.image_shape or img.dtype != np.uint8:
            raise IOError(f'Wrong image: shape {img.shape} dtype {img.dtype} vs {self.image_shape} uint8')
        np.copyto(out, img[:, :, ::-1] if self._xflip[idx] else img)

    def depth_into(self, idx, out):
        d = self.get_depth(idx)
        np.copyto(out, d)

    def __getitem__(self, idx):
        image = np.empty(self.image_shape, dtype=np.uint8)
        self.image_into(idx, image)
        return {'image': image, 'label': self.get_label(idx), 'camera_angles': self.get_camera_angles(idx),
                'depth': self.get_depth(idx).copy() if self._use_depth else np.array([[0]], dtype=np.int32),
                'embedding': self.get_embedding(idx)}

    def get_label(self, idx):
        lab = self._raw_labels()[self._raw_idx[idx]]
        if lab.dtype == np.int64:
            one = np.zeros(self.label_shape, dtype=np.float32)
            one[lab] = 1
            return one
        return lab.copy()

    def get_embedding(self, idx):
        rows, emb = self._raw_embeddings()
        return np.array(emb[rows[self._raw_idx[idx]]]).copy()

    def get_camera_angles(self, idx):
        a = self._raw_camera_angles()[self._raw_idx[idx]].copy()
        if self._xflip[idx]:                 # a mirrored image is seen from the yaw reflected about the mean yaw (:160-162)
            m = self.mean_camera_params[0]
            a[0] = -(a[0] - m) + m
        return a

    def get_depth(self, idx):
        assert self._use_depth
        d = self._load_raw_depth(self._raw_idx[idx])
        assert list(d.shape) == [1, *self.image_shape[1:]] and d.dtype == np.int32, f'Wrong depth: {d.shape} {d.dtype}'
        return d[:, :, ::-1] if self._xflip[idx] else d

    def get_details(self, idx):
        from ..dnnlib import EasyDict
        return EasyDict(raw_idx=int(self._raw_idx[idx]), xflip=(int(self._xflip[idx]) != 0), raw_label=self.get_label(idx))

    def compute_num_classes(self):
        return len(np.unique(self._raw_labels()))

    def close(self):
        self._arc.close()

    # ---- properties the training loop reads ----------------------------------------------------------------------------------
    name = property(lambda self: self._name)
    image_shape = property(lambda self: list(self._raw_shape[1:]))
    num_channels = property(lambda self: self.image_shape[0])
    has_labels = property(lambda self: any(x != 0 for x in self.label_shape))
    has_onehot_labels = property(lambda self: self._raw_labels().dtype == np.int64)
    has_depth = property(lambda self: self.get_depth(0).size > 1)

    @property
    def resolution(self):
        assert self.image_shape[1] == self.image_shape[2]
        return self.image_shape[1]

    @property
    def label_shape(self):
        if self._label_shape is None:
            lab = self._raw_labels()
            self._label_shape = [int(np.max(lab)) + 1] if lab.dtype == np.int64 else list(lab.shape[1:])
        return list(self._label_shape)

    @property
    def label_dim(self):
        assert len(self.label_shape) == 1
        return self.label_shape[0]

    @property
    def mean_camera_params(self):
        """[yaw, pitch, roll, fov, radius] means (dataset.py:229-238): dataset statistics for `dist: custom`, the prior's mean otherwise."""
        if self._mean_cam is None:
            cam = self.cfg.camera
            if cam.origin.angles.dist == 'custom':
                # the statistic is taken over the UNMIRRORED items only (`range(len(raw angles))`), where get_camera_angles never recurses
                ang = np.array([self.get_camera_angles(i) for i in range(len(self._raw_camera_angles()))]).mean(axis=0)
            else:
                ang = get_mean_angles_values(cam.origin.angles)
            self._mean_cam = np.concatenate([ang, np.array([get_mean_sampling_value(cam.fov), get_mean_sampling_value(cam.origin.radius)])])
        return self._mean_cam


def infinite_order(n, rank=0, num_replicas=1, shuffle=True, seed=0, window_size=0.5):
    """The index sequence of the reference's InfiniteSampler (misc.py:112-143): one global stream over a shuffled order that keeps being locally
    re-shuffled inside a sliding window; rank r takes every num_replicas-th element.  Every rank advances the SAME random stream, so the shards stay
    disjoint without communication."""
    assert n > 0 and num_replicas > 0 and 0 <= rank < num_replicas and 0 <= window_size <= 1
    order = np.arange(n)
    rnd, window = None, 0
    if shuffle:
        rnd = np.random.RandomState(seed)
        rnd.shuffle(order)
        window = int(np.rint(order.size * window_size))
    t = 0
    while True:
        i = t % order.size
        if t % num_replicas == rank:
            yield int(order[i])
        if window >= 2:
            j = (i - rnd.randint(window)) % order.size
            order[i], order[j] = order[j], order[i]
        t += 1


class BatchStream:
    """Infinite iterator of PINNED host batches {image u8 [B,C,H,W], label f32 [B,L], camera_angles f32 [B,3], depth i32 [B,1,h,w], embedding f32 [B,E]}.

    A ring of `depth` preallocated pinned buffers; `workers` threads decode the items of a batch directly into its rows, one batch ahead of the
    consumer.  Batch b's buffer is refilled when batch b + depth - 1 is requested, so wrap the stream in `PrefetchLoader(stream, device, depth=d)` with
    depth >= d + 1: PrefetchLoader waits on the host for a batch's copy before handing it out, i.e. before the request that recycles its buffer."""

    def __init__(self, dataset, batch, rank=0, num_replicas=1, seed=0, shuffle=True, workers=8, depth=4, pin=None):
        self.ds, self.batch = dataset, int(batch)
        self.order = infinite_order(len(dataset), rank, num_replicas, shuffle, seed)
        self.pool = ThreadPoolExecutor(max_workers=max(int(workers), 1))
        pin = torch.cuda.is_available() if pin is None else pin
        first = dataset[0]
        C, H, W = dataset.image_shape
        dshape = list(first['depth'].shape)

        def buf(shape, dtype):
            t = torch.empty([self.batch] + list(shape), dtype=dtype)
            return t.pin_memory() if pin else t
        self.ring = [dict(image=buf([C, H, W], torch.uint8), label=buf(first['label'].shape, torch.float32), camera_angles=buf([3], torch.float32),
                          depth=buf(dshape, torch.int32), embedding=buf(first['embedding'].shape, torch.float32)) for _ in range(max(int(depth), 2))]
        self.views = [{k: v.numpy() for k, v in b.items()} for b in self.ring]
        self.turn = 0
        self.pending = self._submit()

    def _fill(self, views, row, idx):
        ds = self.ds
        ds.image_into(idx, views['image'][row])
        views['label'][row] = ds.get_label(idx)
        views['camera_angles'][row] = ds.get_camera_angles(idx)
        if ds._use_depth:
            ds.depth_into(idx, views['depth'][row])
        else:
            views['depth'][row] = 0
        views['embedding'][row] = ds.get_embedding(idx)

    def _submit(self):
        slot = self.turn % len(self.ring)
        self.turn += 1
        idxs = [next(self.order) for _ in range(self.batch)]
        futs = [self.pool.submit(self._fill, self.views[slot], r, i) for r, i in enumerate(idxs)]
        return slot, idxs, futs

    def __iter__(self):
        return self

    def __next__(self):
        slot, idxs, futs = self.pending
        for f in futs:
            f.result()                       # re-raises decode errors
        self.pending = self._submit()        # the next batch decodes while this one is copied / consumed
        out = dict(self.ring[slot])
        self.last_indices = idxs
        return out

    def close(self):
        self.pool.shutdown(wait=True, cancel_futures=True)


def device_inputs(batch):
    """Device-side normalisation of a copied batch (training_loop.py:300-304): image -> float32 in [-1, 1], depth -> float32 in [-1, 1)."""
    from ..dnnlib import EasyDict
    return EasyDict(img=batch['image'].to(torch.float32) / 127.5 - 1.0, c=batch['label'].to(torch.float32), camera_angles=batch['camera_angles'],
                    depth=batch['depth'].to(torch.float32) / 65536 * 2.0 - 1.0, embs=batch['embedding'].to(torch.float32))
