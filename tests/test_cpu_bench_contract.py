"""bench.py contract on a machine without a GPU: the `--impl reference` arm prints ONE JSON line with the agreed keys (executed CPU steps of the oracle
port, bounded `--small` sample here), and the product arm refuses to run without CUDA instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def _run(args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(['--impl', 'reference', '--steps', '1', '--warmup', '0', '--small'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config',
              'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['unit'] == 'images/s' and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['steps'] == 1 and d['warmup'] == 0 and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and 'executed' in d['cpu_baseline']['sample']
    assert d['e2e'] == dict(value=d['value'], unit=d['unit'], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert 'workload' in d['config'] and 'model' not in d['config']
    # executed steps: the reported time per step is consistent with a run that really took that long (no extrapolation)
    assert d['ms_per_step'] * d['steps'] < 600e3


@pytest.mark.skipif(torch.cuda.is_available(), reason='needs a machine without CUDA')
def test_product_arm_fails_loudly_without_a_gpu():
    r = _run(['--steps', '1', '--warmup', '0'], timeout=300)
    assert r.returncode != 0
    assert 'no CUDA device' in (r.stderr + r.stdout)


def test_a_failing_side_leg_is_reported_under_its_key_and_does_not_raise(capsys):
    """gpu_baseline / ginfer / cpu_baseline are comparison legs: an exception inside one becomes {'error': ...} so the headline line is still printed."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_under_test', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    assert bench._side_leg('ok', lambda a, b=0: a + b, 2, b=3) == 5
    out = bench._side_leg('gpu_baseline', lambda: (_ for _ in ()).throw(RuntimeError('CUDNN_STATUS_ALLOC_FAILED')))
    assert out == {'error': 'RuntimeError: CUDNN_STATUS_ALLOC_FAILED'}
    assert 'side leg gpu_baseline failed' in capsys.readouterr().err
