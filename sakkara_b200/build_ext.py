"""
Build of libskk_b200.so: every .cu under sakkara_b200/csrc is compiled for sm_100a with nvcc
(cross-compiles without a GPU) and linked into one in-tree shared library, so that the built
file travels with the repository snapshot to the GPU box.

    python -m sakkara_b200.build_ext [--force] [--jobs N]
"""
import argparse
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ_DIR = os.path.join(CSRC, 'build')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libskk_b200.so')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v', '-I', CSRC, '-I', INCLUDE]


def find_nvcc() -> str:
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found; libskk_b200.so cannot be built')
    return nvcc


def _sources() -> List[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest() -> str:
    h = hashlib.sha256()
    for path in sorted(os.listdir(CSRC)):
        full = os.path.join(CSRC, path)
        if os.path.isfile(full) and path.endswith(('.cu', '.cuh', '.h')):
            h.update(path.encode())
            h.update(open(full, 'rb').read())
    h.update(open(os.path.join(INCLUDE, 'skk_b200.h'), 'rb').read())
    # flags without the absolute include paths, so that the stamp survives a move of the tree
    h.update(' '.join(f for f in NVCC_FLAGS if not os.path.isabs(f)).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = LIB_PATH + '.sha256'
    return os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == _digest()


def _compile(nvcc: str, src: str) -> str:
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')
    cmd = [nvcc] + NVCC_FLAGS + ['-c', src, '-o', obj]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    with open(obj + '.log', 'w') as f:
        f.write(' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError(f'nvcc failed on {src}:\n{proc.stdout}\n{proc.stderr}')
    return obj


def build(force: bool = False, jobs: int = 0, verbose: bool = True) -> str:
    """Compile and link; returns the path of the shared library."""
    if not force and is_current():
        return LIB_PATH
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    sources = _sources()
    jobs = jobs or min(len(sources), os.cpu_count() or 1)
    if verbose:
        print(f'[sakkara_b200] nvcc: {len(sources)} sources, {jobs} jobs, sm_100a', file=sys.stderr)
    with concurrent.futures.ThreadPoolExecutor(max_workers=jobs) as pool:
        objects = list(pool.map(lambda s: _compile(nvcc, s), sources))
    link = [nvcc, '-shared', '-o', LIB_PATH, '-gencode', 'arch=compute_100a,code=sm_100a'] + objects + ['-ldl']
    proc = subprocess.run(link, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f'link failed:\n{proc.stdout}\n{proc.stderr}')
    with open(LIB_PATH + '.sha256', 'w') as f:
        f.write(_digest())
    return LIB_PATH


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--jobs', type=int, default=0)
    a = ap.parse_args()
    print(build(force=a.force, jobs=a.jobs))
