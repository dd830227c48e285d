"""Minimal stand-ins for the two dnnlib containers the hot-path modules use (reference src/dnnlib/util.py:40-120)."""
from typing import Any

import torch


class EasyDict(dict):
    """dict with attribute access."""

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = value

    def __delattr__(self, name: str) -> None:
        del self[name]

    @staticmethod
    def init_recursively(value):
        if isinstance(value, dict):
            return EasyDict(**{k: EasyDict.init_recursively(v) for k, v in value.items()})
        return value


class TensorGroup(EasyDict):
    """Group of tensors aligned on dim 0 (camera params: angles, fov, radius, look_at)."""

    def __init__(self, **kwargs):
        vals = list(kwargs.values())
        assert all(isinstance(v, (torch.Tensor, TensorGroup)) for v in vals)
        assert all(len(v) == len(vals[0]) for v in vals), {k: tuple(v.shape) for k, v in kwargs.items()}
        super().__init__(**kwargs)
        dict.__setitem__(self, '_length', len(vals[0]))

    def __len__(self):
        return dict.__getitem__(self, '_length')

    def __getitem__(self, item):
        if isinstance(item, str):
            return dict.__getitem__(self, item)
        return TensorGroup(**{k: v[item] for k, v in self.items()})

    def items(self):
        return [(k, v) for k, v in dict.items(self) if k != '_length']

    def keys(self):
        return [k for k, _ in self.items()]

    def values(self):
        return [v for _, v in self.items()]

    shape = property(lambda self: [len(self), None])

    # -- member-wise maps: the group behaves like one tensor whose rows are the concatenated members (util.py:109-175) -------------------
    def _map(self, fn):
        return TensorGroup(**{n: fn(v) for n, v in self.items()})

    def _zip(self, other, fn):
        """fn(member, other's member of the same name) for a group, fn(member, other) for a scalar / tensor."""
        if isinstance(other, TensorGroup):
            return TensorGroup(**{n: fn(v, other[n]) for n, v in self.items()})
        return self._map(lambda v: fn(v, other))

    def __add__(self, other):
        return self._zip(other, lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, other):
        return self._zip(other, lambda a, b: a - b)

    def __mul__(self, other):
        return self._zip(other, lambda a, b: a * b)

    __rmul__ = __mul__

    def __pow__(self, other):
        return self._zip(other, lambda a, b: a ** b)

    def to(self, *a, **k):
        return self._map(lambda v: v.to(*a, **k))

    def float(self):
        return self._map(lambda v: v.float())

    def clone(self):
        return self._map(lambda v: v.clone())

    def repeat_interleave(self, *a, **k):
        return self._map(lambda v: v.repeat_interleave(*a, **k))

    def detach(self):
        return self._map(lambda v: v.detach())

    def cpu(self):
        return self._map(lambda v: v.cpu())

    def clamp(self, *a, **k):
        return self._map(lambda v: v.clamp(*a, **k))

    def permute(self, *a, **k):
        return self._map(lambda v: v.permute(*a, **k))

    def mean(self, *a, **k):
        """Member-wise mean (e.g. the mean camera of a sample, inference_utils.py:204); the scalar over everything is `reduce_mean()`."""
        return self._map(lambda v: v.mean(*a, **k))

    def reshape_each(self, new_shape_of):
        """Each member reshaped to `new_shape_of(member)` (inference_utils.py:97)."""
        return self._map(lambda v: v.reshape(new_shape_of(v)))

    @property
    def device(self):
        return self.values()[0].device

    @property
    def shapes(self):
        return [v.shape for v in self.values()]

    @staticmethod
    def cat(groups, dim=0):
        """Groups with the same members joined member by member."""
        names = groups[0].keys()
        assert all(set(g.keys()) == set(names) for g in groups), [g.keys() for g in groups]
        return TensorGroup(**{n: torch.cat([g[n] for g in groups], dim=dim) for n in names})

    def split(self, group_size):
        """Consecutive row blocks of `group_size` (the last one may be short): how training_loop.py:306-319 cuts a batch into per-GPU micro-batches."""
        return [self[i:i + group_size] for i in range(0, len(self), group_size)]

    # -- whole-group reductions (used to tie every member into a loss term: `x + 0.0 * group.max()`, loss.py:171,204,228) ------------------
    def max(self):
        return torch.stack([v.max() for v in self.values()]).max()

    def sum(self):
        return torch.stack([v.sum() for v in self.values()]).sum()

    def numel(self):
        return sum(v.numel() for v in self.values())

    def reduce_mean(self):
        return self.sum() / self.numel()
