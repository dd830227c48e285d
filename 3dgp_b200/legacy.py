"""Reading the reference's network snapshots (`network-snapshot-*.pkl`, written by src/training/training_loop.py:478-484) WITHOUT the reference's
source tree: every network in such a pickle is a `persistent_class` object (src/torch_utils/persistence.py:99-131) that reduces to
`_reconstruct_persistent_obj(meta)` with meta = {type, version, module_src, class_name, state}; `state` is the module's `__dict__`
(`_parameters`, `_buffers`, `_modules` -- themselves persistent objects -- plus `_init_args` / `_init_kwargs`).  The unpickler below intercepts that
hook and every class outside torch / numpy / builtins (OmegaConf nodes, the reference's EasyDict, ...), keeps their state as plain records, flattens the
module tree into a `state_dict` and rebuilds THIS package's Generator / Discriminator from the recorded constructor arguments.  The reference's own
route (`exec` of the embedded module source) is deliberately not taken: the point of a snapshot here is weights + configuration."""
import collections
import io
import pickle

import torch

from .dnnlib import EasyDict


class _Record:
    """State of an object whose class is not importable here."""
    def __init__(self, *args, **kwargs):
        self._args, self._kwargs, self._state = args, kwargs, None

    def __setstate__(self, state):
        self._state = state

    def __reduce_ex__(self, protocol):          # records are read-only
        raise pickle.PicklingError('snapshot records cannot be re-pickled')


def _record_class(module, name):
    return type(name, (_Record,), {'__module__': module, '_origin': f'{module}.{name}'})


class PersistentStub:
    """One persistent object of the snapshot: `class_name`, `state` (its __dict__)."""
    def __init__(self, meta):
        self.class_name = meta['class_name']
        self.state = meta['state'] or {}


# Globals a snapshot legitimately needs (tensors, containers, numpy scalars); everything else -- including builtins such as eval / getattr and torch
# functions outside this list -- unpickles as an inert record.  NOTE: tensor storages are restored by torch.storage._load_from_bytes, i.e. by torch.load on
# an embedded byte string: a snapshot is as trustworthy as a torch.load of the same file (the reference reads it with plain pickle.load).
_ALLOWED = {
    'collections': {'OrderedDict', 'defaultdict'},
    'copyreg': {'_reconstructor'},
    '_codecs': {'encode'},
    'builtins': {'set', 'frozenset', 'slice', 'range', 'complex', 'bytearray', 'bytes', 'object', 'dict', 'list', 'tuple', 'int', 'float', 'str', 'bool'},
    'torch._utils': {'_rebuild_tensor', '_rebuild_tensor_v2', '_rebuild_parameter', '_rebuild_parameter_with_state'},
    'torch.storage': {'_load_from_bytes'},
    'torch': {'Size', 'device', 'FloatStorage', 'HalfStorage', 'BFloat16Storage', 'DoubleStorage', 'LongStorage', 'IntStorage', 'ShortStorage', 'CharStorage',
              'ByteStorage', 'BoolStorage', 'float32', 'float16', 'bfloat16', 'float64', 'int64', 'int32', 'int16', 'int8', 'uint8', 'bool'},
    'numpy': {'ndarray', 'dtype'},
    'numpy.core.multiarray': {'_reconstruct', 'scalar'},
    'numpy._core.multiarray': {'_reconstruct', 'scalar'},
}


class _SnapshotUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if name == '_reconstruct_persistent_obj' and module.endswith('persistence'):
            return PersistentStub
        if name in _ALLOWED.get(module, ()):
            return super().find_class(module, name)
        if module.startswith('torch.nn.modules.'):            # plain torch containers / layers inside the module tree (ModuleList, Sequential, Embedding ...)
            cls = super().find_class(module, name)
            if isinstance(cls, type) and issubclass(cls, torch.nn.Module):
                return cls
        if name == 'EasyDict':
            return EasyDict
        return _record_class(module, name)


def plain(node):
    """OmegaConf containers / nodes, EasyDicts and records -> plain python (dict / list / scalars).  OmegaConf pickles a DictConfig / ListConfig as its
    __dict__ with the children under `_content` and a ValueNode with its payload under `_val`."""
    if isinstance(node, _Record):
        st = node._state if isinstance(node._state, dict) else {}
        if '_content' in st:
            return plain(st['_content'])
        if '_val' in st:
            return plain(st['_val'])
        if isinstance(node._state, dict):
            return {k: plain(v) for k, v in st.items() if not k.startswith('_')}
        return None
    if isinstance(node, dict):
        return {k: plain(v) for k, v in node.items()}
    if isinstance(node, (list, tuple)):
        return [plain(v) for v in node]
    return node


def _module_state(node):
    """__dict__ of a module-like snapshot node: persistent object, record of a non-persistent module class, or a plain torch container."""
    if isinstance(node, PersistentStub):
        return node.state
    if isinstance(node, _Record):
        return node._state if isinstance(node._state, dict) and '_modules' in node._state else None
    if isinstance(node, torch.nn.Module):
        return node.__dict__
    return None


def module_state_dict(node, prefix=''):
    """Flattens a snapshot module tree into an ordered {dotted name: tensor} (parameters and persistent buffers, torch.nn.Module.state_dict order)."""
    st = _module_state(node)
    if st is None:
        raise RuntimeError(f'snapshot entry {prefix or "<root>"} is not a module ({getattr(node, "_origin", type(node).__name__)})')
    out = collections.OrderedDict()
    skip = st.get('_non_persistent_buffers_set', set()) or set()
    for k, v in (st.get('_parameters') or {}).items():
        if v is not None:
            out[prefix + k] = v.detach()
    for k, v in (st.get('_buffers') or {}).items():
        if v is not None and k not in skip:
            out[prefix + k] = v.detach()
    for k, m in (st.get('_modules') or {}).items():
        if m is not None:
            out.update(module_state_dict(m, prefix + k + '.'))
    return out


def _class_name(node):
    return node.class_name if isinstance(node, PersistentStub) else type(node).__name__


def _init_arguments(node):
    st = _module_state(node) or {}
    return plain(list(st.get('_init_args') or [])), plain(st.get('_init_kwargs') or {})


def read_snapshot(f):
    """f: path (optionally .gz) or binary file object.  Returns {name: PersistentStub | record | plain data}: 'G', 'D', 'G_ema', 'training_set_kwargs', ..."""
    if isinstance(f, (str, bytes)):
        import gzip
        opener = gzip.open if str(f).endswith('.gz') else open
        with opener(f, 'rb') as fh:
            return _SnapshotUnpickler(fh).load()
    return _SnapshotUnpickler(f).load()


def _build(node, device, init=None):
    from .training.networks_epigraf import Generator
    from .training.networks_discriminator import Discriminator
    name = _class_name(node)
    cls = {'Generator': Generator, 'Discriminator': Discriminator}.get(name)
    if cls is None:
        raise RuntimeError(f'snapshot holds a {name}; only Generator / Discriminator of the 3dgp model are built here')
    args, kw = _init_arguments(node)
    if init is not None:                   # classes the reference does not decorate carry no constructor record: the caller supplies it
        args, kw = init
    kw = {k: (EasyDict.init_recursively(v) if isinstance(v, dict) else v) for k, v in kw.items()}
    net = cls(*args, **kw)
    sd = module_state_dict(node)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    if missing or unexpected:
        raise RuntimeError(f'snapshot / module mismatch for {name}: missing {list(missing)[:5]}, unexpected {list(unexpected)[:5]}')
    return net.eval().requires_grad_(False).to(device)


_OWN_FORMAT = '3dgp_b200.snapshot.v1'


def _to_plain_config(v):
    """EasyDict / TensorGroup-free copy of constructor arguments: nested dicts, lists and scalars only."""
    if isinstance(v, dict):
        return {k: _to_plain_config(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_to_plain_config(x) for x in v]
    return v


def save_network_pkl(f, networks, init, **entries):
    """The writing side of `training_loop.py:478-484` for THIS package's modules.  networks: {'G': module, 'D': ..., 'G_ema': ...}; init: {name: (args,
    kwargs)} -- the constructor arguments (config.build_networks' call) -- ; entries: plain data stored alongside (training_set_kwargs, cur_nimg, ...).
    A network is stored as {format, class_name, init, state_dict (CPU tensors)}: weights + configuration, no code and no module objects, so the file loads
    through the same allow-listed unpickler as a reference snapshot.  A reference installation reads the weights with
    `G.load_state_dict(pickle.load(f)['G']['state_dict'])` -- the parameter / buffer names are the reference's."""
    data = dict(entries)
    for name, net in networks.items():
        args, kw = init[name]
        data[name] = {'format': _OWN_FORMAT, 'class_name': type(net).__name__, 'init': (_to_plain_config(list(args)), _to_plain_config(dict(kw))),
                      'state_dict': collections.OrderedDict((k, v.detach().cpu().clone()) for k, v in net.state_dict().items())}
    if isinstance(f, (str, bytes)):
        import gzip
        opener = gzip.open if str(f).endswith('.gz') else open
        with opener(f, 'wb') as fh:
            pickle.dump(data, fh, protocol=4)
    else:
        pickle.dump(data, f, protocol=4)


def _build_own(entry, device):
    from .training.networks_epigraf import Generator
    from .training.networks_discriminator import Discriminator
    cls = {'Generator': Generator, 'Discriminator': Discriminator}.get(entry['class_name'])
    if cls is None:
        raise RuntimeError(f'snapshot holds a {entry["class_name"]}; only Generator / Discriminator of the 3dgp model are built here')
    args, kw = entry['init']
    net = cls(*args, **{k: (EasyDict.init_recursively(v) if isinstance(v, dict) else v) for k, v in kw.items()})
    net.load_state_dict(entry['state_dict'], strict=True)
    return net.eval().requires_grad_(False).to(device)


def load_network_pkl(f, device='cpu', names=('G', 'D', 'G_ema'), init=None):
    """The reference's `legacy.load_network_pkl` role: networks of a snapshot as THIS package's modules (eval mode, no grad), other entries as plain data.
    Reads the reference's snapshots and the ones `save_network_pkl` writes.  init: optional {name: (args, kwargs)} constructor arguments for entries
    whose class keeps none."""
    data = read_snapshot(f)
    out = {}
    for k, v in data.items():
        if isinstance(v, dict) and v.get('format') == _OWN_FORMAT:
            if k in names:
                out[k] = _build_own(v, device)
        elif _module_state(v) is not None:
            if k in names:
                out[k] = _build(v, device, (init or {}).get(k))
        else:
            out[k] = plain(v)
    return out
