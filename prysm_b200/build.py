"""Build libprysm_b200.so in-tree with nvcc for sm_100a.

    python prysm_b200/build.py [--force] [--verbose]      (by path: importing the package needs the library this builds)

The shared library lands in prysm_b200/_lib/ (git-ignored, but it travels to the GPU box
with the gpurun snapshot).  No torch extension machinery: the boundary is a plain C ABI.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, '_lib')
LIB = os.path.join(LIBDIR, 'libprysm_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
    hsh = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ['../../include/prysm_b200.h']:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            hsh.update(f.encode())
            hsh.update(open(p, 'rb').read())
    hsh.update(' '.join(FLAGS).encode())
    return hsh.hexdigest()


def _header_digest():
    hsh = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ['../../include/prysm_b200.h']:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p) and not f.endswith('.cu'):
            hsh.update(f.encode())
            hsh.update(open(p, 'rb').read())
    hsh.update(' '.join(FLAGS).encode())
    return hsh.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, 'build.stamp')
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    objdir = os.path.join(LIBDIR, 'obj')
    os.makedirs(objdir, exist_ok=True)
    hdig = _header_digest()

    def compile_one(src):
        # incremental: an object is rebuilt when its source, any header or the flags changed
        obj = os.path.join(objdir, src[:-3] + '.o')
        ostamp = obj + '.stamp'
        odig = hashlib.sha256(open(os.path.join(CSRC, src), 'rb').read() + hdig.encode()).hexdigest()
        if not force and os.path.exists(obj) and os.path.exists(ostamp) and open(ostamp).read() == odig:
            return obj
        cmd = [NVCC, *FLAGS, '-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{r.stdout}\n{r.stderr}')
        if verbose and r.stderr:
            print(r.stderr)
        open(ostamp, 'w').write(odig)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [NVCC, '-shared', '-o', LIB, *objs, '-lcuda', '-ldl']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    open(stamp, 'w').write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
