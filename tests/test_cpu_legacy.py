"""Reference snapshot pickles load into THIS package's modules without the reference's source tree (3dgp_b200/legacy.py).  The fixture
tests/golden/snapshot_small.pkl.gz was written by the unmodified reference's persistence machinery (src/torch_utils/persistence.py:99-131, the
dict layout of training_loop.py:478-484) for the small golden networks, with deterministic low-entropy weights (oracle/cases.py::snapshot_fill) and the
embedded module sources replaced by a placeholder (oracle/make_golden.py::gen_snapshot).
Parity note: the fixture's configuration objects are the harness's EasyDicts; a production snapshot carries OmegaConf DictConfig nodes, whose
`_content` / `_val` layout `legacy.plain` unwraps -- that branch is exercised with hand-built records below, not with a real OmegaConf pickle
(OmegaConf is not installed here): unpinned for that one conversion."""
import collections
import importlib
import os

import numpy as np
import torch

from conftest import ROOT
from oracle import cases

FIXTURE = os.path.join(ROOT, 'tests', 'golden', 'snapshot_small.pkl.gz')


def test_snapshot_networks_load_with_the_reference_weights_and_constructor_arguments():
    lg = importlib.import_module('3dgp_b200.legacy')
    cfgm = importlib.import_module('3dgp_b200.config')
    kw = {k: v for k, v in cases.net_kwargs('small').items() if k != 'learn_camera_dist'}
    cfg = cfgm.make_config(**kw)
    # the reference does not decorate `Discriminator` itself (networks_discriminator.py:201): its constructor record is not in the pickle
    d_init = ([], dict(cfg=cfg.model.discriminator, input_resolution=cfg.training.patch.resolution, img_channels=4, block_kwargs=dict(freeze_layers=0),
                       mapping_kwargs={}, epilogue_kwargs=dict(mbstd_group_size=4, feat_predict_dim=cfg.dataset.embedding_dim), num_fp16_res=0, conv_clamp=None))
    nets = lg.load_network_pkl(FIXTURE, init={'D': d_init})
    assert set(nets) >= {'G', 'D', 'G_ema', 'training_set_kwargs'}
    assert nets['training_set_kwargs'] == dict(path='synthetic', resolution=kw['img_resolution'], use_labels=True)
    G, D, Ge = nets['G'], nets['D'], nets['G_ema']
    assert type(G).__module__.startswith('3dgp_b200.') and type(D).__module__.startswith('3dgp_b200.')
    assert not G.training and not any(p.requires_grad for p in G.parameters())
    # constructor arguments came out of the pickle
    assert G.img_resolution == kw['img_resolution'] and G.z_dim == kw['z_dim'] and G.cfg.tri_plane.res == kw['tri_res'] and G.cfg.cmax == kw['cmax']
    for tag, net in (('G.', G), ('D.', D), ('G.', Ge)):
        sd = net.state_dict()
        assert len(sd) > 20
        for k, v in sd.items():
            assert np.array_equal(v.numpy(), cases.snapshot_fill(tag + k, tuple(v.shape))), tag + k


def test_plain_unwraps_omegaconf_style_records():
    """DictConfig / ListConfig pickle as their __dict__ with the children under `_content`, value nodes with the payload under `_val`."""
    lg = importlib.import_module('3dgp_b200.legacy')

    def rec(state):
        r = lg._record_class('omegaconf.dictconfig', 'DictConfig')()
        r.__setstate__(state)
        return r
    node = rec({'_metadata': object(), '_parent': None, '_content': {'res': rec({'_val': 512, '_metadata': None}), 'mlp': rec({'_content': {'hid_dim': rec({'_val': 64})}}),
                                                                  'betas': rec({'_content': [rec({'_val': 0.0}), rec({'_val': 0.99})]})}})
    assert lg.plain(node) == {'res': 512, 'mlp': {'hid_dim': 64}, 'betas': [0.0, 0.99]}


def test_snapshot_reader_refuses_arbitrary_globals():
    """The unpickler resolves only an explicit allow-list of globals (tensor rebuild hooks, containers, numpy scalars, torch.nn module classes): anything
    else -- os / posixpath functions, builtins such as eval or getattr, other torch functions -- becomes an inert record class instead of being imported."""
    import io
    import pickle
    lg = importlib.import_module('3dgp_b200.legacy')
    for obj in (os.path.join, eval, getattr, torch.load, torch.hub.load):
        out = lg._SnapshotUnpickler(io.BytesIO(pickle.dumps({'x': obj}))).load()
        assert isinstance(out['x'], type) and issubclass(out['x'], lg._Record), obj
    ok = lg._SnapshotUnpickler(io.BytesIO(pickle.dumps({'t': torch.arange(3.0), 'd': collections.OrderedDict(a=1), 's': {1, 2}}))).load()
    assert torch.equal(ok['t'], torch.arange(3.0)) and ok['d'] == {'a': 1} and ok['s'] == {1, 2}


def test_own_snapshot_round_trip_and_reference_side_state_dict(tmp_path):
    """save_network_pkl -> load_network_pkl: same constructor arguments, same weights, eval mode / no grad like the reference's snapshots; the file holds no
    module objects (it loads through the allow-listed unpickler), and its `state_dict` entries carry the reference's parameter names (a reference
    installation reads them with load_state_dict)."""
    import json
    import pickle
    lg = importlib.import_module('3dgp_b200.legacy')
    cfgm = importlib.import_module('3dgp_b200.config')
    kw = {k: v for k, v in cases.net_kwargs('small').items() if k != 'learn_camera_dist'}
    cfg = cfgm.make_config(**kw, learn_camera_dist=True)
    torch.manual_seed(3)
    G, D = cfgm.build_networks(cfg, 'cpu', fp32_D=True)
    g = cfg.model.generator
    init = {'G': ([], dict(cfg=g, img_resolution=cfg.dataset.resolution, img_channels=3, mapping_kwargs=dict(camera_cond=False, camera_cond_drop_p=0.0, mean_camera_params=None),
                           fused_modconv_default='inference_only', num_fp16_res=0, conv_clamp=None)),
            'D': ([], dict(cfg=cfg.model.discriminator, input_resolution=cfg.training.patch.resolution, img_channels=4, block_kwargs=dict(freeze_layers=0), mapping_kwargs={},
                           epilogue_kwargs=dict(mbstd_group_size=4, feat_predict_dim=cfg.dataset.embedding_dim), num_fp16_res=0, conv_clamp=None))}
    init['G_ema'] = init['G']
    path = str(tmp_path / 'network-snapshot-000000.pkl')
    lg.save_network_pkl(path, dict(G=G, D=D, G_ema=G), init, training_set_kwargs=dict(path='x.zip', resolution=32), cur_nimg=12345)
    back = lg.load_network_pkl(path)
    assert back['cur_nimg'] == 12345 and back['training_set_kwargs'] == dict(path='x.zip', resolution=32)
    for name, net in (('G', G), ('D', D), ('G_ema', G)):
        sd, sb = net.state_dict(), back[name].state_dict()
        assert list(sd) == list(sb) and all(torch.equal(sd[k], sb[k]) for k in sd)
        assert not back[name].training and not any(p.requires_grad for p in back[name].parameters())
    assert back['G'].synthesis.camera_adaptor is not None                     # learn_camera_dist=true survived through the stored configuration
    raw = pickle.load(open(path, 'rb'))                                          # plain pickle: no class of this package is referenced by the file
    assert raw['G']['class_name'] == 'Generator' and isinstance(raw['G']['init'][1]['cfg'], dict) and type(raw['G']['init'][1]['cfg']) is dict
    meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'networks_meta.json')))
    assert set(meta['D_keys']) == set(raw['D']['state_dict'])                    # the reference Discriminator's own state-dict keys (golden meta was written by the reference)
