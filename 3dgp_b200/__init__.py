"""3dgp_b200 -- Blackwell (sm_100a) implementation of 3DGP's per-image hot path.

The directory name starts with a digit, so import it with
    import importlib; gp = importlib.import_module('3dgp_b200')
Sub-packages mirror the reference layout (src/torch_utils/ops/*, src/training/*), see INTEGRATION.md.
"""
from . import _lib  # noqa: F401  (ctypes binding of lib3dgp_b200.so; loading is lazy)

__all__ = ['_lib']
