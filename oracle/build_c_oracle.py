"""Compiles oracle/logp_c.c with gcc into oracle/_build/liboracle_glm.so (test infrastructure)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'logp_c.c')
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'liboracle_glm.so')


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ['gcc', '-O3', '-march=x86-64-v2', '-fopenmp', '-fPIC', '-shared', '-o', LIB, SRC, '-lm']
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build(force=True))
