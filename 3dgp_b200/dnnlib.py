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

    def max(self):
        return torch.stack([v.max() for v in self.values()]).max()

    def to(self, *a, **k):
        return TensorGroup(**{n: v.to(*a, **k) for n, v in self.items()})
