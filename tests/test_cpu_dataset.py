"""Training-set reader (3dgp_b200/training/dataset.py) against what the UNMODIFIED reference reads from the same files
(tests/golden/dataset_golden.npz, written by oracle/make_dataset_golden.py through src/training/dataset.py::ImageFolderDataset and
src/torch_utils/misc.py::InfiniteSampler).  Byte / index work: every comparison is bit-exact."""
import importlib
import io
import json
import os
import zipfile

import numpy as np
import pytest
import torch

from conftest import ROOT

dsmod = importlib.import_module('3dgp_b200.training.dataset')
dn = importlib.import_module('3dgp_b200.dnnlib')
GOLD = os.path.join(ROOT, 'tests', 'golden')
ZIP = os.path.join(GOLD, 'tiny_dataset.zip')


def _cfg(mirror, dist='uniform', c_dim=3):
    return dn.EasyDict.init_recursively(dict(
        c_dim=c_dim, use_embeddings=False, mirror=mirror,
        camera=dict(fov=dict(dist='uniform', min=10.0, max=45.0),
                    origin=dict(radius=dict(dist='normal', mean=1.0, std=0.0),
                                angles=dict(dist=dist, yaw=dict(min=-1.57, max=1.57, mean=0.0, std=0.4),
                                            pitch=dict(min=0.785398163, max=2.35619449, mean=1.57, std=0.2))))))


VARIANTS = {'plain': dict(cfg=_cfg(False)), 'mirror': dict(cfg=_cfg(True)), 'subset': dict(cfg=_cfg(True), max_size=7, random_seed=3),
            'custom': dict(cfg=_cfg(True, dist='custom'))}


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLD, 'dataset_golden.npz'))


@pytest.mark.parametrize('tag', sorted(VARIANTS))
def test_items_equal_the_reference_bit_for_bit(gold, tag):
    ds = dsmod.ImageFolderDataset(path=ZIP, resolution=16, use_depth=False, **VARIANTS[tag])
    assert len(ds) == int(gold[f'{tag}/len'])
    assert ds.image_shape == list(gold[f'{tag}/image_shape']) and ds.label_shape == list(gold[f'{tag}/label_shape'])
    assert ds.resolution == 16 and ds.num_channels == 3 and ds.has_labels and ds.has_onehot_labels and ds.label_dim == 3
    assert ds.compute_num_classes() == int(gold[f'{tag}/num_classes']) and ds.name == 'tiny_dataset'
    np.testing.assert_array_equal(np.asarray(ds.mean_camera_params, dtype=np.float64), gold[f'{tag}/mean_camera_params'])
    items = [ds[i] for i in range(len(ds))]
    for k in ('image', 'label', 'camera_angles', 'depth', 'embedding'):
        got = np.stack([it[k] for it in items])
        assert got.dtype == gold[f'{tag}/{k}'].dtype and got.shape == gold[f'{tag}/{k}'].shape, k
        np.testing.assert_array_equal(got, gold[f'{tag}/{k}'], err_msg=k)
    np.testing.assert_array_equal([ds.get_details(i).raw_idx for i in range(len(ds))], gold[f'{tag}/raw_idx'])
    np.testing.assert_array_equal([ds.get_details(i).xflip for i in range(len(ds))], gold[f'{tag}/xflip'])
    ds.close()


def test_sampler_order_equals_the_reference(gold):
    n = int(gold['mirror/len'])
    keys = [k for k in gold.files if k.startswith('sampler/')]
    assert len(keys) == 4
    for k in keys:
        rank, rep, seed, shuffle = [int(v) for v in k.split('/')[1].split('_')]
        it = dsmod.infinite_order(n, rank=rank, num_replicas=rep, shuffle=bool(shuffle), seed=seed)
        np.testing.assert_array_equal([next(it) for _ in range(100)], gold[k], err_msg=k)


def test_rank_shards_are_disjoint_and_cover_the_stream():
    n, rep = 24, 4
    whole = dsmod.infinite_order(n, 0, 1, True, 9)
    ref = [next(whole) for _ in range(200)]
    shards = [dsmod.infinite_order(n, r, rep, True, 9) for r in range(rep)]
    inter = [next(shards[t % rep]) for t in range(200)]
    assert inter == ref                                            # rank r sees elements r, r + R, ... of the one global stream


def test_directory_layout_and_missing_metadata(tmp_path):
    with zipfile.ZipFile(ZIP) as z:
        z.extractall(tmp_path / 'tiny_dataset')
    a = dsmod.ImageFolderDataset(path=ZIP, cfg=_cfg(False))
    b = dsmod.ImageFolderDataset(path=str(tmp_path / 'tiny_dataset'), cfg=_cfg(False))
    assert len(a) == len(b) == 12
    for i in (0, 5, 11):
        np.testing.assert_array_equal(a[i]['image'], b[i]['image']); np.testing.assert_array_equal(a[i]['label'], b[i]['label'])
    os.remove(tmp_path / 'tiny_dataset' / 'dataset.json')
    c = dsmod.ImageFolderDataset(path=str(tmp_path / 'tiny_dataset'), cfg=_cfg(False, c_dim=0))
    assert not c.has_labels and c[0]['label'].shape == (0,) and np.all(c[0]['camera_angles'] == 0)
    with pytest.raises(AssertionError):                            # labels requested, none stored (dataset.py:66-68)
        dsmod.ImageFolderDataset(path=str(tmp_path / 'tiny_dataset'), cfg=_cfg(False)).get_label(0)
    with pytest.raises(IOError):
        dsmod.ImageFolderDataset(path=ZIP, resolution=32, cfg=_cfg(False))
    with pytest.raises(IOError):
        dsmod.ImageFolderDataset(path=str(tmp_path / 'nothing.txt'), cfg=_cfg(False))


def test_depth_maps_8_and_16_bit_and_mirror(tmp_path):
    import PIL.Image
    root = tmp_path / 'dd'; os.makedirs(root)
    rs = np.random.RandomState(1)
    d16 = rs.randint(0, 65536, size=(8, 8)).astype(np.uint16); d8 = rs.randint(0, 256, size=(8, 8)).astype(np.uint8)
    for name, d in (('a', d16), ('b', d8)):
        PIL.Image.fromarray(rs.randint(0, 256, size=(8, 8, 3)).astype(np.uint8), 'RGB').save(root / f'{name}.png')
        PIL.Image.fromarray(d).save(root / f'{name}_depth.png')
    ds = dsmod.ImageFolderDataset(path=str(root), use_depth=True, cfg=_cfg(True, c_dim=0))
    assert len(ds) == 4 and ds.has_depth                            # the depth files are not counted as images; mirror doubles the set
    np.testing.assert_array_equal(ds[0]['depth'], d16.astype(np.int32)[None])
    np.testing.assert_array_equal(ds[1]['depth'], (d8.astype(np.int32) * 256)[None])           # 8-bit maps are scaled to the 16-bit range (dataset.py:319-320)
    np.testing.assert_array_equal(ds[2]['depth'], d16.astype(np.int32)[None][:, :, ::-1])      # mirrored copy
    assert ds[0]['depth'].dtype == np.int32


def test_depth_items_equal_the_reference_bit_for_bit():
    """use_depth=True on the committed depth fixture (16-bit grey, 8-bit grey, 8-bit RGB depth maps, all five PNG scanline filters) against the items the
    UNMODIFIED reference class returned for it; the reference's `pyspng.load` was stood in for by the specification decoder oracle/png_spec.py
    (pyspng is absent from this image), the product decodes with PIL: two independent decoders, one reference logic (dataset.py:164-173, 310-323)."""
    g = np.load(os.path.join(GOLD, 'dataset_depth_golden.npz'))
    ds = dsmod.ImageFolderDataset(path=os.path.join(GOLD, 'tiny_dataset_depth.zip'), resolution=16, use_depth=True, cfg=_cfg(True, c_dim=2))
    assert len(ds) == int(g['len']) == 12 and bool(ds.has_depth) == bool(g['has_depth'])
    items = [ds[i] for i in range(len(ds))]
    for k in ('image', 'label', 'depth'):
        got = np.stack([it[k] for it in items])
        assert got.dtype == g[k].dtype and got.shape == g[k].shape, k
        np.testing.assert_array_equal(got, g[k], err_msg=k)
    assert g['depth'].dtype == np.int32 and g['depth'].max() > 255 * 128      # the fixture really spans the 16-bit range
    ds.close()


def test_specification_png_decoder_agrees_with_pil_on_every_filter_type():
    """The stand-in for pyspng (oracle/png_spec.py, written from the PNG specification) and PIL decode each other's files identically."""
    import PIL.Image
    from oracle import png_spec
    rs = np.random.RandomState(5)
    for shape, dt in (((16, 16, 3), np.uint8), ((16, 16), np.uint8), ((16, 16), np.uint16), ((5, 7, 4), np.uint8)):
        a = (rs.randint(0, 65536, size=shape) if dt == np.uint16 else rs.randint(0, 256, size=shape)).astype(dt)
        for ft in range(5):
            data = png_spec.save(a, ft)
            assert png_spec.load(data).dtype == dt and np.array_equal(png_spec.load(data), a)
            assert np.array_equal(np.array(PIL.Image.open(io.BytesIO(data))).astype(dt), a), (shape, ft)
        b = io.BytesIO(); PIL.Image.fromarray(a).save(b, format='png', compress_level=9)     # PIL chooses its own (adaptive) filters
        assert np.array_equal(png_spec.load(b.getvalue()), a)


def test_batch_stream_fills_pinned_style_buffers_in_sampler_order():
    ds = dsmod.ImageFolderDataset(path=ZIP, cfg=_cfg(True))
    stream = dsmod.BatchStream(ds, batch=5, rank=1, num_replicas=2, seed=4, workers=3, depth=3, pin=False)
    order = dsmod.infinite_order(len(ds), 1, 2, True, 4)
    seen = []
    for _ in range(7):                                              # more batches than ring slots: buffers are reused
        b = next(stream)
        idx = [next(order) for _ in range(5)]
        assert stream.last_indices == idx
        assert b['image'].dtype == torch.uint8 and tuple(b['image'].shape) == (5, 3, 16, 16) and b['depth'].dtype == torch.int32
        for r, i in enumerate(idx):
            it = ds[i]
            np.testing.assert_array_equal(b['image'][r].numpy(), it['image']); np.testing.assert_array_equal(b['label'][r].numpy(), it['label'])
            np.testing.assert_array_equal(b['camera_angles'][r].numpy(), it['camera_angles'])
        seen.append(b['image'].clone())
    assert not torch.equal(seen[0], seen[3])                        # slot 0 was refilled with a different batch
    x = dsmod.device_inputs(b)
    assert x.img.dtype == torch.float32 and float(x.img.min()) >= -1.0 and float(x.img.max()) <= 1.0
    np.testing.assert_array_equal(x.img.numpy(), b['image'].numpy().astype(np.float32) / 127.5 - 1.0)     # training_loop.py:300
    stream.close(); ds.close()


def test_decode_errors_surface_to_the_consumer(tmp_path):
    root = tmp_path / 'bad'; os.makedirs(root)
    import PIL.Image
    PIL.Image.fromarray(np.zeros((8, 8, 3), np.uint8), 'RGB').save(root / 'a.png')
    PIL.Image.fromarray(np.zeros((4, 4, 3), np.uint8), 'RGB').save(root / 'b.png')      # wrong size
    ds = dsmod.ImageFolderDataset(path=str(root), cfg=_cfg(False, c_dim=0))
    stream = dsmod.BatchStream(ds, batch=2, shuffle=False, workers=2, pin=False)
    with pytest.raises(IOError):
        next(stream)
    stream.close()


def test_embeddings_memmap_and_generator_side_batch(tmp_path):
    """Pre-extracted feature embeddings (the knowledge-distillation targets, dataset.py:355-362) follow the file -> row table of their description;
    `training_loop.sample_gen_batch` draws the generator's conditioning from the training set (training_loop.py:305-319)."""
    from util import write_training_set
    tl = importlib.import_module('3dgp_b200.training.training_loop')
    root = str(tmp_path / 'set')
    extra = write_training_set(root, n=10, res=8, c_dim=4, depth=True, emb_dim=6, seed=3)
    emb, rows = extra.pop('_embeddings'), extra.pop('_rows')
    cfg = _cfg(True, dist='custom', c_dim=4); cfg.update(extra)
    ds = dsmod.ImageFolderDataset(path=root, use_depth=True, cfg=cfg)
    assert len(ds) == 20 and ds.has_depth
    for i in (0, 3, 13):
        raw = int(ds.get_details(i).raw_idx)
        np.testing.assert_array_equal(ds[i]['embedding'], emb[rows[ds._image_fnames[raw]]])
    full = dn.EasyDict.init_recursively(dict(camera=dict(cfg.camera, look_at=dict(radius=dict(dist='uniform', min=0.0, max=0.2),
                                             angles=dict(dist='spherical_uniform', yaw=dict(min=-3.14, max=3.14), pitch=dict(min=0.0, max=3.14))))))
    rng = np.random.RandomState(5)
    g = tl.sample_gen_batch(full, ds, 8, torch.device('cpu'), z_dim=16, gpc_spoof_p=0.5, rng=rng)
    pick = [np.random.RandomState(5).randint(len(ds))]              # first draw of the same stream
    assert tuple(g.z.shape) == (8, 16) and tuple(g.c.shape) == (8, 4) and float(g.c.sum()) == 8.0
    np.testing.assert_array_equal(g.c[0].numpy(), ds.get_label(pick[0]))
    np.testing.assert_allclose(g.camera_params.angles[0].numpy(), ds.get_camera_angles(pick[0]), rtol=0, atol=0)      # `custom`: cameras sit on dataset angles
    assert tuple(g.camera_params.look_at.shape) == (8, 3) and tuple(g.camera_angles_cond.shape) == (8, 3)


def test_our_dataset_under_the_reference_loader_sampler_and_grid_setup():
    """The reference's OWN data path around the dataset class (src/train.py:109-134 `init_dataset`, training_loop.py:96-100 InfiniteSampler + DataLoader,
    inference_utils.py:20-39 `setup_snapshot_image_grid`) run once on the reference's ImageFolderDataset and once on this repo's, on the committed depth
    fixture: constructor keywords as train.py passes them, `mean_camera_params`, `get_camera_angles`, collated batches (every key, dtype and value), the
    snapshot grid -- identical; this repo's class also survives DataLoader worker processes."""
    import copy
    import types
    from oracle import png_spec, ref_harness as rh
    if not rh.available():
        pytest.skip('the unmodified reference is only present in the build container')
    ns = rh.load()
    import src.training.dataset as ref_ds
    import src.training.inference_utils as iu
    from src.torch_utils import misc
    ED = ns.dnnlib.EasyDict
    dcfg = ED.init_recursively(json.loads(json.dumps(_cfg(True, dist='custom', c_dim=2))))
    cam_cfg = copy.deepcopy(rh.make_cfg()[0]['camera']); cam_cfg['origin']['angles']['dist'] = 'custom'
    cfg = ED.init_recursively(dict(training=dict(use_depth=True), camera=cam_cfg))
    path = os.path.join(GOLD, 'tiny_dataset_depth.zip')

    def sampler(ds):        # torch >= 2.2 dropped Sampler.__init__(data_source), which misc.py:118 still calls: set the fields, run the unmodified __iter__
        s = misc.InfiniteSampler.__new__(misc.InfiniteSampler)
        s.dataset, s.rank, s.num_replicas, s.shuffle, s.seed, s.window_size = ds, 0, 1, True, 0, 0.5
        return s

    def walk(cls):
        ds = cls(path=path, max_size=None, use_depth=True, cfg=dcfg)                                                      # train.py:111-114
        ds = cls(path=path, use_depth=True, cfg=dcfg, resolution=ds.resolution, max_size=len(ds), random_seed=0)            # train.py:115-116,180
        out = dict(mean=torch.from_numpy(ds.mean_camera_params), angles=np.array([ds.get_camera_angles(i) for i in range(len(ds))]),
                   flags=(ds.has_labels, ds.has_depth, list(ds.image_shape), list(ds.label_shape), ds.resolution, ds.num_channels))
        it = iter(torch.utils.data.DataLoader(dataset=ds, sampler=sampler(ds), batch_size=4, num_workers=0))
        out['batches'] = [next(it), next(it)]
        np.random.seed(0); torch.manual_seed(0)
        out['grid'] = iu.setup_snapshot_image_grid(training_set=ds, cfg=cfg)
        return out, ds

    old = ref_ds.pyspng
    ref_ds.pyspng = types.SimpleNamespace(load=png_spec.load)      # pyspng is absent here: the specification decoder stands in (oracle/png_spec.py)
    try:
        a, _ = walk(ref_ds.ImageFolderDataset)
    finally:
        ref_ds.pyspng = old
    b, ours = walk(dsmod.ImageFolderDataset)
    assert torch.equal(a['mean'], b['mean']) and np.array_equal(a['angles'], b['angles']) and a['flags'] == b['flags']
    for x, y in zip(a['batches'], b['batches']):
        assert set(x) == set(y) == {'image', 'label', 'camera_angles', 'depth', 'embedding'}
        assert all(x[k].dtype == y[k].dtype and torch.equal(x[k], y[k]) for k in x)
    (ga, ia, da, la, ca), (gb, ib, db, lb, cb) = a['grid'], b['grid']
    assert tuple(ga) == tuple(gb) and np.array_equal(ia, ib) and np.array_equal(da, db) and np.array_equal(la, lb) and all(torch.equal(ca[k], cb[k]) for k in ca.keys())
    it = iter(torch.utils.data.DataLoader(dataset=ours, sampler=sampler(ours), batch_size=4, num_workers=2, prefetch_factor=2))      # training_loop.py:100, c.data_loader_kwargs
    w = next(it)
    assert all(torch.equal(w[k], b['batches'][0][k]) for k in w)
    del it
