#!/usr/bin/env python
"""
Small evaluations that reach every path of the row kernel and the finish / exchange kernels, for
compute-sanitizer:

    compute-sanitizer --tool racecheck python tools/sanitize_run.py
    compute-sanitizer --tool memcheck  python tools/sanitize_run.py

Shapes: crossed Poisson with primary groups much longer than a tile (hand-pipelined common path + two-group
tiles), with short groups (tiles of three or more groups), nested logistic (an element fed by many slots),
linear regression with a batch tile of 4 and the chain kernel, and the crossed shape again through the
sharded pipeline with this rank as its only peer.  Every result is compared with the CPU oracle.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sakkara_b200 import synth                                  # noqa: E402
from sakkara_b200.valuegrad import ValueGradFunction            # noqa: E402


def run(name, plan, dtype, batch=1, peer_self=False):
    from oracle.logp_numpy import parity_error
    f = ValueGradFunction(plan, dtype=dtype, max_batch=batch, device=0)
    if peer_self:
        f.skk.peer_init([f.skk.peer_alloc(1)], 0, 1)
    theta = synth.start_point(plan, batch, seed=3, dtype=dtype)
    logp, grad = f(theta)
    logp, grad = np.atleast_1d(logp), np.atleast_2d(grad)
    tol = 1e-4 if dtype == 'float32' else 1e-10
    worst = 0.0
    for b in range(min(batch, 2)):
        e_l, e_g = parity_error(logp[b], grad[b], plan, theta[b].astype(np.float64))
        worst = max(worst, e_l, e_g)
    again = f(theta)
    same = np.array_equal(np.atleast_1d(again[0]), logp) and np.array_equal(np.atleast_2d(again[1]), grad)
    print(f'[sanitize] {name:34s} {dtype} batch={batch} rows={plan.n_rows} err={worst:.2e} reproducible={same}', flush=True)
    assert worst <= tol and same, name
    f.close()


def main():
    scale = int(os.environ.get('SKK_SANITIZE_ROWS', '120000'))
    cols, _ = synth.poisson_columns(scale, 12, 40, seed=5)          # 10k rows per primary group: 18 tiles each
    run('crossed, long groups', synth.poisson_plan(cols), 'float32')
    run('crossed, long groups, peer self', synth.poisson_plan(cols), 'float32', peer_self=True)
    run('crossed, long groups, fp64', synth.poisson_plan(cols), 'float64')
    cols, _ = synth.poisson_columns(scale // 4, 400, 30, seed=6)    # 75 rows per group: many groups per tile
    run('crossed, short groups', synth.poisson_plan(cols), 'float32')
    cols, _ = synth.logistic_columns(scale, 8, 10, seed=7)
    run('nested logistic', synth.logistic_plan(cols), 'float32')
    run('nested logistic, peer self', synth.logistic_plan(cols), 'float32', peer_self=True)
    cols, _ = synth.linreg_columns(scale // 2, 50, seed=8)
    run('linear regression, batch 4', synth.linreg_plan(cols), 'float64', batch=4)
    run('linear regression, chains 32', synth.linreg_plan(cols), 'float64', batch=32)
    print('[sanitize] all shapes ok')


if __name__ == '__main__':
    main()
