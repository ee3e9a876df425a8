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


def _includes(path: str, seen=None) -> List[str]:
    """Transitive closure of the quoted #include files of a source (dependency tracking)."""
    seen = seen if seen is not None else []
    for line in open(path):
        line = line.strip()
        if line.startswith('#include "'):
            name = line.split('"')[1]
            full = os.path.normpath(os.path.join(os.path.dirname(path), name))
            if os.path.exists(full) and full not in seen:
                seen.append(full)
                _includes(full, seen)
    return seen


def _flags_digest() -> str:
    # flags without the absolute include paths, so that stamps survive a move of the tree
    return ' '.join(f for f in NVCC_FLAGS if not os.path.isabs(f))


def _source_digest(src: str) -> str:
    h = hashlib.sha256()
    for path in [src] + sorted(_includes(src)):
        h.update(os.path.basename(path).encode())
        h.update(open(path, 'rb').read())
    h.update(_flags_digest().encode())
    return h.hexdigest()


def _digest() -> str:
    h = hashlib.sha256()
    for src in _sources():
        h.update(_source_digest(src).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = LIB_PATH + '.sha256'
    return os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == _digest()


def _compile(nvcc: str, src: str) -> str:
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')
    digest = _source_digest(src)
    if os.path.exists(obj) and os.path.exists(obj + '.sha256') and open(obj + '.sha256').read() == digest:
        return obj
    cmd = [nvcc] + NVCC_FLAGS + ['-c', src, '-o', obj]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    with open(obj + '.log', 'w') as f:
        f.write(' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError(f'nvcc failed on {src}:\n{proc.stdout}\n{proc.stderr}')
    with open(obj + '.sha256', 'w') as f:
        f.write(digest)
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
