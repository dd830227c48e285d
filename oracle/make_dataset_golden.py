"""TEST INFRASTRUCTURE ONLY (build container): writes the tiny training-set fixture and what the UNMODIFIED reference reads from it.

    python oracle/make_dataset_golden.py

  tests/golden/tiny_dataset.zip          12 RGB PNGs (16 x 16) in two class folders + dataset.json (labels, camera angles), the layout dataset_tool.py writes
  tests/golden/dataset_golden.npz        items of src/training/dataset.py::ImageFolderDataset (mirror on / off, max_size subset, custom-angle statistics) and
                                         index sequences of src/torch_utils/misc.py::InfiniteSampler for several (rank, replicas, seed)
  tests/golden/tiny_dataset_depth.zip    6 RGB PNGs + `<name>_depth.png` maps: 16-bit greyscale, 8-bit greyscale, 8-bit RGB (first channel is the depth), written
                                         with all five PNG scanline filters by oracle/png_spec.py
  tests/golden/dataset_depth_golden.npz  items (image, depth) of the same reference class with use_depth=True, mirror on
The reference decodes depth maps with `pyspng.load` (dataset.py:314); pyspng is not installed here, so for the depth golden a stand-in `pyspng` module whose
`load` is oracle/png_spec.py::load (a decoder written from the PNG specification, no imaging library) is injected before the UNMODIFIED reference module is
imported: reference logic + an independent decoder.  RGB images of the depth fixture then go through the stand-in too (dataset.py:301-302)."""
import io
import json
import os
import sys
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
ZIP = os.path.join(GOLD, 'tiny_dataset.zip')


def dataset_cfg(mirror, dist='uniform', c_dim=3):
    return dict(c_dim=c_dim, use_embeddings=False, mirror=mirror,
                camera=dict(fov=dict(dist='uniform', min=10.0, max=45.0),
                            origin=dict(radius=dict(dist='normal', mean=1.0, std=0.0),
                                        angles=dict(dist=dist, yaw=dict(min=-1.57, max=1.57, mean=0.0, std=0.4),
                                                    pitch=dict(min=0.785398163, max=2.35619449, mean=1.57, std=0.2)))))


def write_fixture():
    import PIL.Image
    rs = np.random.RandomState(20240)
    labels, angles = [], []
    with zipfile.ZipFile(ZIP, 'w', zipfile.ZIP_STORED) as z:
        for i in range(12):
            name = f'{i % 3:05d}/img{i:08d}.png'
            img = rs.randint(0, 256, size=(16, 16, 3)).astype(np.uint8)
            b = io.BytesIO(); PIL.Image.fromarray(img, 'RGB').save(b, format='png', compress_level=0, optimize=False)
            z.writestr(zipfile.ZipInfo(name, date_time=(2020, 1, 1, 0, 0, 0)), b.getvalue())
            labels.append([name, int(i % 3)])
            angles.append([name, [float(rs.uniform(-1.5, 1.5)), float(rs.uniform(0.8, 2.3)), 0.0]])
        z.writestr(zipfile.ZipInfo('dataset.json', date_time=(2020, 1, 1, 0, 0, 0)), json.dumps(dict(labels=labels, camera_angles=angles)))


DEPTH_ZIP = os.path.join(GOLD, 'tiny_dataset_depth.zip')


def write_depth_fixture():
    from oracle import png_spec
    rs = np.random.RandomState(777)
    labels = []
    with zipfile.ZipFile(DEPTH_ZIP, 'w', zipfile.ZIP_STORED) as z:
        def put(name, data):
            z.writestr(zipfile.ZipInfo(name, date_time=(2020, 1, 1, 0, 0, 0)), data)
        for i in range(6):
            base = f'{i % 2:05d}/img{i:08d}'
            put(base + '.png', png_spec.save(rs.randint(0, 256, size=(16, 16, 3)).astype(np.uint8), filter_type=i % 5))
            ramp = (np.add.outer(np.arange(16), np.arange(16)) * 1500 + rs.randint(0, 1500, size=(16, 16)))      # smooth + noise, spans the 16-bit range
            if i % 3 == 0:
                depth = ramp.astype(np.uint16)                                    # 16-bit greyscale
            elif i % 3 == 1:
                depth = (ramp >> 8).astype(np.uint8)                              # 8-bit greyscale (the reader scales it by 256)
            else:
                depth = np.stack([(ramp >> 8).astype(np.uint8), rs.randint(0, 256, (16, 16)).astype(np.uint8), np.zeros((16, 16), np.uint8)], axis=2)   # RGB: channel 0
            put(base + '_depth.png', png_spec.save(depth, filter_type=(i + 2) % 5))
            labels.append([base + '.png', int(i % 2)])
        put('dataset.json', json.dumps(dict(labels=labels)))


def depth_golden(ns):
    """Runs the unmodified reference reader with use_depth=True; `pyspng` is the specification decoder (see the module docstring)."""
    import types
    from oracle import png_spec
    import src.training.dataset as ref_ds
    assert ref_ds.pyspng is None, 'pyspng is installed after all: drop the stand-in and regenerate'
    ref_ds.pyspng = types.SimpleNamespace(load=png_spec.load)
    try:
        cfg = ns.dnnlib.EasyDict.init_recursively(dataset_cfg(True, c_dim=2))
        ds = ref_ds.ImageFolderDataset(path=DEPTH_ZIP, resolution=16, use_depth=True, cfg=cfg)
        items = [ds[i] for i in range(len(ds))]
        out = {'len': np.int64(len(ds)), 'has_depth': np.bool_(ds.has_depth)}
        for k in ('image', 'label', 'depth'):
            out[k] = np.stack([it[k] for it in items])
        ds.close()
    finally:
        ref_ds.pyspng = None
    np.savez_compressed(os.path.join(GOLD, 'dataset_depth_golden.npz'), **out)
    print('wrote', DEPTH_ZIP, os.path.getsize(DEPTH_ZIP), 'bytes; depth', out['depth'].shape, out['depth'].dtype, int(out['depth'].min()), int(out['depth'].max()))


def main():
    os.makedirs(GOLD, exist_ok=True)
    write_fixture()
    write_depth_fixture()
    ns = ref_harness.load()
    sys.path.insert(0, ref_harness.REF_ROOT)
    from src.training.dataset import ImageFolderDataset
    from src.torch_utils.misc import InfiniteSampler
    out = {}
    variants = {'plain': dict(cfg=dataset_cfg(False)), 'mirror': dict(cfg=dataset_cfg(True)), 'subset': dict(cfg=dataset_cfg(True), max_size=7, random_seed=3),
                'custom': dict(cfg=dataset_cfg(True, dist='custom'))}
    for tag, kw in variants.items():
        kw = dict(kw); kw['cfg'] = ns.dnnlib.EasyDict.init_recursively(kw['cfg'])
        ds = ImageFolderDataset(path=ZIP, resolution=16, use_depth=False, **kw)
        out[f'{tag}/len'] = np.int64(len(ds))
        out[f'{tag}/image_shape'] = np.array(ds.image_shape)
        out[f'{tag}/label_shape'] = np.array(ds.label_shape)
        out[f'{tag}/mean_camera_params'] = np.asarray(ds.mean_camera_params, dtype=np.float64)
        out[f'{tag}/num_classes'] = np.int64(ds.compute_num_classes())
        items = [ds[i] for i in range(len(ds))]
        for k in ('image', 'label', 'camera_angles', 'depth', 'embedding'):
            out[f'{tag}/{k}'] = np.stack([it[k] for it in items])
        out[f'{tag}/raw_idx'] = np.array([ds.get_details(i).raw_idx for i in range(len(ds))])
        out[f'{tag}/xflip'] = np.array([ds.get_details(i).xflip for i in range(len(ds))])
        if tag == 'mirror':
            for (rank, rep, seed, shuffle) in [(0, 1, 0, True), (1, 2, 0, True), (3, 4, 7, True), (0, 2, 5, False)]:
                # torch >= 2.2 dropped Sampler.__init__(data_source), which the reference's constructor still calls (misc.py:118): set the fields it
                # would set and run its UNMODIFIED __iter__
                smp = InfiniteSampler.__new__(InfiniteSampler)
                smp.dataset, smp.rank, smp.num_replicas, smp.shuffle, smp.seed, smp.window_size = ds, rank, rep, shuffle, seed, 0.5
                it = iter(smp)
                out[f'sampler/{rank}_{rep}_{seed}_{int(shuffle)}'] = np.array([int(next(it)) for _ in range(100)])
        ds.close()
    np.savez_compressed(os.path.join(GOLD, 'dataset_golden.npz'), **out)
    depth_golden(ns)
    print('wrote', ZIP, os.path.getsize(ZIP), 'bytes;', len(out), 'golden arrays')


if __name__ == '__main__':
    main()
