"""The reference's composed experiment configuration loads unchanged: tests/golden/reference_experiment_config.json is what src/infra/launch.py:24-81
saves for the README's ImageNet-256 command (composed from the reference's configs/ tree by oracle/compose_config.py, committed as a fixture because
/root/reference does not travel).  `yaml.safe_load -> EasyDict.init_recursively` of that file must build the same networks as the package defaults."""
import importlib
import json
import os

import torch

from conftest import ROOT


def _ref_cfg():
    dn = importlib.import_module('3dgp_b200.dnnlib')
    return dn.EasyDict.init_recursively(json.load(open(os.path.join(ROOT, 'tests', 'golden', 'reference_experiment_config.json'))))


def _leaves(d, prefix=''):
    for k, v in d.items():
        if isinstance(v, dict):
            yield from _leaves(v, prefix + k + '.')
        else:
            yield prefix + k, v


def test_package_defaults_equal_the_reference_composition():
    """Every key of 3dgp_b200.config.make_config() that the reference's composed config also defines carries the reference's value
    (configs/{camera,model,training,dataset}/*.yaml + README.md:57 overrides); keys the modules read are all present in the reference's file."""
    cfgm = importlib.import_module('3dgp_b200.config')
    ours = dict(_leaves(cfgm.make_config(learn_camera_dist=True)))
    ref = dict(_leaves(_ref_cfg()))
    missing = [k for k in ours if k not in ref]
    # keys the port adds on top of the reference's file (defaults the reference takes from function signatures)
    allowed_extra = {'model.generator.density_bias', 'model.discriminator.hyper_mod'}
    assert not [k for k in missing if k not in allowed_extra and not k.startswith('model.generator.depth_adaptor.camera.')
                and not k.startswith('model.generator.camera_adaptor.camera.')], missing
    diff = {}
    for k, v in ours.items():
        if k in ref:
            r = ref[k]
            same = (abs(float(v) - float(r)) <= 1e-6 * max(1.0, abs(float(r)))) if isinstance(v, (int, float)) and isinstance(r, (int, float)) and not isinstance(v, bool) else (v == r)
            if not same:
                diff[k] = (v, r)
    # two deliberate departures, both stated: BASELINE.json configs[1] quotes the metric at 48 samples/ray (configs/model/3dgp.yaml: 32), and the
    # package keeps gamma = 'auto' (train.py:173's heuristic, which bench.py evaluates) where the README command pins 0.05 -- a loss scalar, no shape
    assert diff.pop('model.generator.num_ray_steps', (48, 32)) == (48, 32)
    assert diff.pop('model.loss_kwargs.gamma', ('auto', 0.05)) == ('auto', 0.05)
    assert not diff, diff


def test_networks_build_from_the_reference_experiment_config():
    """Generator / Discriminator constructed from the reference's own configuration file: same parameter names, shapes and counts as from the package
    defaults (105.7 M / 143.1 M parameters at cmax=1024: SURVEY.md 8d's all-reduce sizes)."""
    cfgm = importlib.import_module('3dgp_b200.config')
    Gr, Dr = cfgm.build_networks(_ref_cfg(), 'cpu')
    Go, Do = cfgm.build_networks(cfgm.make_config(learn_camera_dist=True), 'cpu')
    for a, b in ((Gr, Go), (Dr, Do)):
        sa, sb = {k: tuple(v.shape) for k, v in a.state_dict().items()}, {k: tuple(v.shape) for k, v in b.state_dict().items()}
        assert sa == sb
    ng = sum(p.numel() for p in Gr.parameters()); nd = sum(p.numel() for p in Dr.parameters())
    assert abs(ng - 105.72e6) < 0.6e6 and abs(nd - 143.06e6) < 0.6e6, (ng, nd)
