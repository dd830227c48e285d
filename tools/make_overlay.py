#!/usr/bin/env python
"""Lays this repo's hot path over a checkout of snap-research/3dgp (as a command):

    python tools/make_overlay.py /path/to/3dgp            # writes into /path/to/3dgp/src/...
    python tools/make_overlay.py /path/to/3dgp --dest /tmp/3dgp_b200_overlay    # copies the checkout's src/ there first, leaves the original alone

What it writes under `src/` (everything else stays the reference's: dnnlib, rendering_utils, training_utils, loss, training_loop, metrics, ...):
    _lib.py, build.py, csrc/, lib3dgp_b200.so (when built)         the C-ABI loader and the CUDA sources;  ../include/gp3d_b200.h next to src/
    torch_utils/custom_ops.py                                       get_plugin -> the sm_100a plugin objects
    torch_utils/ops/{bias_act,upfirdn2d,filtered_lrelu,fma,conv2d_gradfix,conv2d_resample,tc,modconv,raymarch}.py
    training/{layers,networks_epigraf,networks_stylegan2,networks_discriminator,networks_depth_adaptor,networks_camera_adaptor,tri_plane_renderer}.py

Two mechanical edits are applied to the copied Python files, because of how the reference pickles networks (src/torch_utils/persistence.py:99-131 stores the
defining module's SOURCE in every snapshot and `exec`s it in an anonymous module when the installed code differs):
  * relative imports (`from ..dnnlib import X`) become the absolute form the reference uses (`from src.dnnlib import X`) -- a relative import cannot be
    resolved inside an exec'd anonymous module;
  * the classes the reference decorates with `@persistence.persistent_class` get the decorator back, so `training_loop.py:478-484` writes
    self-describing snapshots of these modules and `legacy.load_network_pkl` / `scripts/utils.py` read them."""
import argparse
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(HERE, '3dgp_b200')

FILES = {
    '': ['_lib.py', 'build.py'],
    'torch_utils': ['custom_ops.py'],
    'torch_utils/ops': ['bias_act.py', 'upfirdn2d.py', 'filtered_lrelu.py', 'fma.py', 'conv2d_gradfix.py', 'conv2d_resample.py', 'tc.py', 'modconv.py', 'raymarch.py'],
    'training': ['layers.py', 'networks_epigraf.py', 'networks_stylegan2.py', 'networks_discriminator.py', 'networks_depth_adaptor.py', 'networks_camera_adaptor.py',
                 'tri_plane_renderer.py'],
}
# classes the reference marks persistent (src/training/*.py, `@persistence.persistent_class`); Discriminator itself is not (networks_discriminator.py:201)
PERSISTENT = {
    'training/layers.py': ['FullyConnectedLayer', 'MappingNetwork', 'Conv2dLayer', 'ScalarEncoder1d', 'FourierEncoder1d'],
    'training/networks_epigraf.py': ['TriPlaneMLP', 'SynthesisBlocksSequence', 'SynthesisNetwork', 'Generator'],
    'training/networks_stylegan2.py': ['SynthesisLayer', 'ToRGBLayer', 'SynthesisBlock', 'SynthesisNetwork', 'Generator'],
    'training/networks_discriminator.py': ['DiscriminatorBlock', 'MinibatchStdLayer', 'DiscriminatorEpilogue'],
    'training/networks_depth_adaptor.py': ['DepthAdaptor'],
    'training/networks_camera_adaptor.py': ['ParamsAdaptor', 'CameraAdaptor'],
    'training/tri_plane_renderer.py': ['ImportanceRenderer'],
}
_REL = re.compile(r'^(?P<ind>\s*)from (?P<dots>\.+)(?P<mod>[\w\.]*) import ', re.M)


def absolutise(text, package):
    """`from ..a.b import c` inside `package` (e.g. 'src.training') -> `from src.a.b import c`; `from . import x` -> `from src.training import x`."""
    parts = package.split('.')

    def fix(m):
        up = len(m.group('dots')) - 1
        base = parts[:len(parts) - up] if up else parts
        assert base, f'relative import climbs above the package root: {m.group(0)!r}'
        target = '.'.join(base + ([m.group('mod')] if m.group('mod') else []))
        return f"{m.group('ind')}from {target} import "
    return _REL.sub(fix, text)


def decorate(text, names):
    for n in names:
        pat = re.compile(rf'^class {n}\(', re.M)
        assert len(pat.findall(text)) == 1, n
        text = pat.sub(f'@persistence.persistent_class\nclass {n}(', text)
    # the import goes after the module docstring / before the first import statement
    first = re.search(r'^(import |from )', text, re.M)
    return text[:first.start()] + 'from src.torch_utils import persistence\n' + text[first.start():]


def install(ref_root, dest=None, root_package='src'):
    src = os.path.join(ref_root, 'src')
    assert os.path.isdir(os.path.join(src, 'training')), f'{ref_root} does not look like a 3dgp checkout (no src/training)'
    if dest is not None:
        os.makedirs(dest, exist_ok=True)
        shutil.copytree(src, os.path.join(dest, 'src'), dirs_exist_ok=True, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        out = os.path.join(dest, 'src')
    else:
        dest, out = ref_root, src
    written = []
    for sub, names in FILES.items():
        for n in names:
            rel = f'{sub}/{n}' if sub else n
            text = open(os.path.join(PKG, sub, n)).read()
            text = absolutise(text, '.'.join([root_package] + [p for p in sub.split('/') if p]))
            if rel in PERSISTENT:
                text = decorate(text, PERSISTENT[rel])
            with open(os.path.join(out, sub, n), 'w') as f:
                f.write(text)
            written.append(rel)
    shutil.copytree(os.path.join(PKG, 'csrc'), os.path.join(out, 'csrc'), dirs_exist_ok=True)
    os.makedirs(os.path.join(dest, 'include'), exist_ok=True)
    shutil.copyfile(os.path.join(HERE, 'include', 'gp3d_b200.h'), os.path.join(dest, 'include', 'gp3d_b200.h'))
    lib = os.path.join(PKG, 'lib3dgp_b200.so')
    if os.path.exists(lib):
        shutil.copyfile(lib, os.path.join(out, 'lib3dgp_b200.so'))
    return out, written


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('reference', help='root of a snap-research/3dgp checkout (the directory that holds src/)')
    ap.add_argument('--dest', default=None, help='write a patched COPY of src/ here instead of patching the checkout in place')
    a = ap.parse_args()
    out, written = install(a.reference, a.dest)
    print(f'{len(written)} files laid over {out}:')
    for w in written:
        print('  ', w)
    print('build the library with:  python -c "import sys; sys.path.insert(0, %r); from src import build; build.build()"' % os.path.dirname(out))


if __name__ == '__main__':
    sys.exit(main())
