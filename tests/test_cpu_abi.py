"""CPU suite: the C-ABI shared library loads and exports every symbol include/gp3d_b200.h declares; argument
validation paths that need no GPU."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, 'include', 'gp3d_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(gp3d_[a-z0-9_]+)\s*\(', hdr)))


def test_every_declared_symbol_is_exported_and_bound(gp):
    L = gp._lib.lib()
    syms = _declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(L, s), f'{s} declared in include/gp3d_b200.h but not exported by lib3dgp_b200.so'
        assert s in gp._lib.PROTOTYPES, f'{s} has no ctypes prototype in 3dgp_b200/_lib.py'
    assert sorted(gp._lib.PROTOTYPES) == syms


def test_version_and_arch(gp):
    L = gp._lib.lib()
    assert L.gp3d_version() == 1
    assert L.gp3d_built_arch() == 100


def test_out_size_formula_matches_reference_integer_rule(gp):
    L = gp._lib.lib()
    for (n, up, down, p0, p1, fs) in [(16, 2, 1, 2, 1, 4), (17, 1, 1, 1, 1, 4), (16, 1, 2, 1, 1, 4), (7, 3, 2, 2, 3, 5), (1, 2, 1, 2, 1, 4), (64, 1, 1, 30, 30, 61)]:
        assert L.gp3d_upfirdn2d_out_size(n, up, down, p0, p1, fs) == (n * up + p0 + p1 - fs + down) // down


def test_argument_errors_without_gpu(gp):
    L = gp._lib.lib()
    # null pointers / bad enums are rejected before any launch
    rc = L.gp3d_bias_act(None, None, None, None, None, None, 0, 16, 1, 1, 0, 3, 0.2, 1.0, -1.0, None)
    assert rc == -1 and b'non-null' in L.gp3d_last_error()
    rc = L.gp3d_upfirdn2d(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 0, 1, 1, 4, 4, 16, 16, 4, 1, 4, 4, 0, 1, 1, 1, 0, 0, 0, 0, 0, 1.0, 1, 1, 1, 1, 1, 1, None)
    assert rc == -1 and b'upsampling factor' in L.gp3d_last_error()


def test_product_has_no_cpu_fallback(gp):
    import importlib
    import pytest
    import torch
    ba = importlib.import_module('3dgp_b200.torch_utils.ops.bias_act')
    up = importlib.import_module('3dgp_b200.torch_utils.ops.upfirdn2d')
    with pytest.raises(RuntimeError):
        ba.bias_act(torch.zeros(2, 3), torch.zeros(3))
    with pytest.raises(RuntimeError):
        up.upfirdn2d(torch.zeros(1, 1, 4, 4), None)
    with pytest.raises(RuntimeError):
        ba.bias_act(torch.zeros(2, 3), impl='ref')


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, '3dgp_b200')
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith('.py'):
                src = open(os.path.join(dp, fn)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f'{fn} imports the oracle'
