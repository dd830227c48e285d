"""TEST INFRASTRUCTURE (oracle side): composes the reference's Hydra configuration tree the way `src/infra/launch.py:24-28` does for the README's
ImageNet command (README.md:57) -- defaults list of configs/config.yaml, `# @package _group_` placement, `${a.b}` interpolation, and
`recursive_instantiate` of the `_target_` helpers of src/infra/utils.py:143-193 -- without Hydra / OmegaConf (not installed here).  The result is what
the launcher saves as `experiment_config.yaml` (launch.py:81) and `src/train.py:149-152` loads.  Written once into tests/golden/ by
oracle/make_golden.py; the product never imports this module.

The infra / env groups (git hashes, slurm, paths) are skipped: nothing on the hot path reads them."""
import copy
import math
import os
import re

import yaml

TARGETS = {
    'src.infra.utils.divide': lambda dividend, divisor: dividend / divisor,                          # utils.py:167
    'src.infra.utils.log2_divide': lambda dividend, divisor: int(math.log2(dividend / divisor)),       # utils.py:172
    'src.infra.utils.product_ab': lambda a, b: a * b,                                                  # utils.py:161
    'src.infra.utils.compute_magnitude_ema_beta': lambda batch_size: 0.5 ** (batch_size / (20 * 1e3)),  # utils.py:182
    'src.infra.utils.linspace': lambda val_from, val_to, num_steps: [val_from + (val_to - val_from) * i / (num_steps - 1) for i in range(num_steps)],
}
_REF = re.compile(r'\$\{([^}]+)\}')


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


def _get(root, path):
    node = root
    for part in path.split('.'):
        node = node[part]
    return node


def _resolve(root, node, depth=0):
    assert depth < 50, 'interpolation cycle'
    if isinstance(node, dict):
        return {k: _resolve(root, v, depth + 1) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(root, v, depth + 1) for v in node]
    if isinstance(node, str):
        m = _REF.fullmatch(node)
        if m:                                       # whole-value reference: the referenced NODE (dict / number / ...)
            if m.group(1).startswith(('env.', 'hydra:', 'env:')):
                return node
            return _resolve(root, _get(root, m.group(1)), depth + 1)
        def sub(mm):
            if mm.group(1).startswith(('env.', 'hydra:', 'env:')):
                return mm.group(0)
            v = _resolve(root, _get(root, mm.group(1)), depth + 1)
            if isinstance(v, dict) and '_target_' in v:
                v = _instantiate(v)
            return str(v)
        return _REF.sub(sub, node)
    return node


def _instantiate(node):
    """src/infra/utils.py:132-140 (recursive_instantiate): dicts carrying `_target_` become the helper's return value."""
    if isinstance(node, dict):
        if '_target_' in node:
            kw = {k: _instantiate(v) for k, v in node.items() if k != '_target_'}
            return TARGETS[node['_target_']](**kw)
        return {k: _instantiate(v) for k, v in node.items()}
    if isinstance(node, list):
        return [_instantiate(v) for v in node]
    return node


def compose(config_dir, overrides=None, skip_groups=('env', 'infra')):
    """overrides: {'model.generator.cmax': 1024, 'dataset': 'imagenet', ...} -- group choices or dotted value overrides (Hydra command line)."""
    overrides = dict(overrides or {})
    top = yaml.safe_load(open(os.path.join(config_dir, 'config.yaml')))
    cfg = {}
    for item in top['defaults']:
        if isinstance(item, str):                   # 'group/base.yaml' or a root-level file
            group = item.split('/')[0] if '/' in item else None
            path = item
        else:
            (group, choice), = item.items()
            choice = overrides.pop(group, choice)
            path = f'{group}/{choice}.yaml'
        if (group or os.path.splitext(path)[0]) in skip_groups:
            continue
        doc = yaml.safe_load(open(os.path.join(config_dir, path))) or {}
        if group is None:
            _merge(cfg, doc)
        else:                                       # '# @package _group_': the file's content lives under its group key
            _merge(cfg.setdefault(group, {}), doc)
    for key, val in overrides.items():              # dotted value overrides
        parts = key.split('.')
        node = cfg
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = val
    return _instantiate(_resolve(cfg, cfg))


README_IMAGENET_OVERRIDES = {   # README.md:57
    'dataset': 'imagenet', 'dataset.resolution': 256, 'model.loss_kwargs.gamma': 0.05, 'training.resume': None,
    'model.generator.cmax': 1024, 'model.discriminator.cmax': 1024, 'model.generator.cbase': 65536, 'model.discriminator.cbase': 65536,
}
