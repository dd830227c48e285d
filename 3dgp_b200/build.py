"""Builds lib3dgp_b200.so (hand-written sm_100a CUDA behind the C ABI of include/gp3d_b200.h) in-tree with nvcc.

No torch C++ headers are involved: the library is a plain C-ABI shared object, loaded with ctypes by
`3dgp_b200/_lib.py`.  nvcc cross-compiles without a GPU, so this runs in the build container; the resulting
.so travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'lib3dgp_b200.so')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC', '-DGP3D_ARCH=100', '-I', INCLUDE, '--expt-relaxed-constexpr',
]
# Memory-bound streaming ops follow the reference's --use_fast_math build (bias_act.py:46, upfirdn2d.py:31);
# the ray-march keeps IEEE expf/log1pf/div for fp32 parity with the CPU oracle.
FAST_MATH = {'bias_act.cu', 'upfirdn2d.cu', 'misc.cu', 'filtered_lrelu.cu'}


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: lib3dgp_b200 cannot be built')


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest(path, flags):
    h = hashlib.sha256()
    h.update(' '.join(flags).encode())
    with open(path, 'rb') as f:
        h.update(f.read())
    for hdr in sorted(os.listdir(CSRC)):
        if hdr.endswith(('.cuh', '.h')):
            with open(os.path.join(CSRC, hdr), 'rb') as f:
                h.update(f.read())
    with open(os.path.join(INCLUDE, 'gp3d_b200.h'), 'rb') as f:
        h.update(f.read())
    return h.hexdigest()


def build(verbose=False, force=False):
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    objs = []
    for src in sources():
        flags = list(NVCC_FLAGS) + (['--use_fast_math'] if src in FAST_MATH else [])
        spath = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, src[:-3] + '.o')
        stamp = obj + '.sha'
        dig = _digest(spath, flags)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        jobs.append((src, [nvcc] + flags + ['-c', spath, '-o', obj], stamp, dig))

    def run(job):
        src, cmd, stamp, dig = job
        if verbose:
            print('[3dgp_b200.build]', ' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        if verbose and r.stderr.strip():
            print(r.stderr)
        with open(stamp, 'w') as f:
            f.write(dig)
        return src

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB) or force:
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcuda']
        if verbose:
            print('[3dgp_b200.build]', ' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return LIB


if __name__ == '__main__':
    print(build(verbose=True, force='--force' in sys.argv))
