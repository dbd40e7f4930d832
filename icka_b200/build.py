"""Build libicka_b200.so in-tree with nvcc for sm_100a (no torch dependency in the library).

    python -m icka_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT_DIR = os.path.join(HERE, 'lib')
BUILD_DIR = os.path.join(HERE, 'lib', 'obj')
LIB = os.path.join(OUT_DIR, 'libicka_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(os.path.dirname(HERE), 'include', 'icka_b200.h'))
    return max(os.path.getmtime(f) for f in files)


def _compile(src):
    obj = os.path.join(BUILD_DIR, src[:-3] + '.o')
    cmd = [NVCC, *FLAGS, '-c', os.path.join(CSRC, src), '-o', obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    return src, obj, p.returncode, p.stdout + p.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD_DIR, exist_ok=True)
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    if not os.path.isfile(NVCC):
        raise RuntimeError(f'nvcc not found at {NVCC}; libicka_b200.so must be built before use')
    objs, log = [], []
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for src, obj, rc, out in ex.map(_compile, sources()):
            log.append(f'==== {src}\n{out}')
            if rc != 0:
                raise RuntimeError(f'nvcc failed for {src}:\n{out}')
            objs.append(obj)
    with open(os.path.join(OUT_DIR, 'build.log'), 'w') as f:
        f.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    cmd = [NVCC, '-shared', '-o', LIB, *objs, '-gencode', 'arch=compute_100a,code=sm_100a', '-cudart', 'static']
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError('link failed:\n' + p.stdout + p.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
