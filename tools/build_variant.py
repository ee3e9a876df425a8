#!/usr/bin/env python
"""
Build a variant of libskk_b200.so with extra -D flags, for A/B timing on the GPU box:

    python tools/build_variant.py k6 -DSKK_PIPE_ROWS=6
    SKK_B200_LIB=sakkara_b200/lib/variants/libskk_b200_k6.so python bench.py ...

Objects go to sakkara_b200/csrc/build/<name>/, the library to sakkara_b200/lib/variants/ (git-ignored,
but shipped to the GPU box with the snapshot).
"""
import concurrent.futures
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sakkara_b200 import build_ext  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    nvcc = build_ext.find_nvcc()
    obj_dir = os.path.join(build_ext.OBJ_DIR, name)
    lib_dir = os.path.join(build_ext.LIB_DIR, 'variants')
    os.makedirs(obj_dir, exist_ok=True)
    os.makedirs(lib_dir, exist_ok=True)
    only = os.environ.get('SKK_VARIANT_ONLY')   # e.g. "poisson_bt1": other objects are taken from the main build

    # With SKK_VARIANT_ONLY=<pattern>[,<pattern>...] (e.g. f32_poisson, or chain_f64,f64_normal) the library
    # holds the kernels of the matching sources only (lookups of the others return null: plans of other families fail at creation), which
    # keeps a variant at ~15 MB instead of ~85 MB.
    stub_names = []

    def compile_one(src):
        base = os.path.basename(src)[:-3]
        if only and base not in ('skk_api', 'skk_microbench') and not any(pat in base for pat in only.split(',')):
            if base.startswith('skk_row_'):
                stub_names.append(('row', base[len('skk_row_'):]))
            else:
                stub_names.append(('chain', base[len('skk_chain_'):]))
            return None
        obj = os.path.join(obj_dir, base + '.o')
        cmd = [nvcc] + build_ext.NVCC_FLAGS + flags + ['-c', src, '-o', obj]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        open(obj + '.log', 'w').write(' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
        if proc.returncode != 0:
            raise RuntimeError(proc.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count()) as pool:
        objects = list(pool.map(compile_one, build_ext._sources()))
    objects = [o for o in objects if o]
    if stub_names:
        stub = os.path.join(obj_dir, 'stubs.cu')
        with open(stub, 'w') as f:
            f.write('namespace skk {\n')
            for kind, nm in stub_names:
                if kind == 'row':
                    f.write(f'const void *row_kernel_{nm}(int, int, int, int) {{ return nullptr; }}\n')
                else:
                    f.write(f'const void *chain_kernel_{nm}(int, int, int) {{ return nullptr; }}\n')
            f.write('}\n')
        subprocess.run([nvcc, '-Xcompiler', '-fPIC', '-c', stub, '-o', stub[:-3] + '.o'], check=True)
        objects.append(stub[:-3] + '.o')
    out = os.path.join(lib_dir, f'libskk_b200_{name}.so')
    subprocess.run([nvcc, '-shared', '-o', out, '-gencode', 'arch=compute_100a,code=sm_100a'] + objects + ['-ldl'],
                   check=True)
    print(out)


if __name__ == '__main__':
    main()
