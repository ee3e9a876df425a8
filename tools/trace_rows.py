#!/usr/bin/env python
"""
Per-warp timeline of the row kernel (skk_trace_enable / skk_trace_read): where the time of a launch goes.

    python tools/trace_rows.py --config c4 [--rows N] [--emulate-world W] [--evals 5]

Prints, over all warps of the grid: kernel span, prologue (entry -> ready for the first tile), wait for the
first tile, tile loop, epilogue; and the loop time of a warp against the tiles it processed, by tile class
(1 group / 2 groups / more), derived on the host from the plan's segment offsets.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='c4')
    ap.add_argument('--rows', type=int, default=0)
    ap.add_argument('--emulate-world', type=int, default=1)
    ap.add_argument('--evals', type=int, default=5)
    ap.add_argument('--out', default='')
    a = ap.parse_args()
    import torch
    import bench
    from sakkara_b200 import synth
    from sakkara_b200.valuegrad import ValueGradFunction
    cfg = bench.CONFIGS[a.config]
    rows_total = a.rows or cfg['rows']
    plan = bench.make_plan(a.config, rows_total, 0, a.emulate_world)
    batch = min(cfg['batch'], 4)
    f = ValueGradFunction(plan, dtype=cfg['dtype'], max_batch=batch, device=0, n_rows_total=rows_total)
    skk = f.skk
    if a.emulate_world > 1:
        os.environ['SKK_FORCE_PEER'] = '1'
        from sakkara_b200.distributed import connect
        connect(skk, 0, 1)
    theta = synth.start_point(plan, batch, dtype=cfg['dtype'])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda:0')
    skk.trace_enable(True)
    st = skk.stats()
    grid, warps, wt = st['grid'], st['block'] // 32, st['rows_per_tile']
    kp = f.kernel_plan
    n_tiles = st['n_tiles']
    runs = []
    for _ in range(a.evals):
        flush.zero_()
        torch.cuda.synchronize()
        f(theta)
        runs.append(skk.trace_read().astype(np.int64))
    fin = skk.trace_read(finish=True).astype(np.int64)
    tr = runs[-1]
    per_warp = tr[:, 5:8]                      # tiles per class, counted by the kernel from the warp's schedule
    cls_total = per_warp.sum(0)
    act = per_warp.sum(1) > 0
    t0, t1, t2, t3, t4 = (tr[:, k] for k in range(5))
    base = t0.min()
    span = t4.max() - base

    def q(x):
        x = x[act] if len(x) == len(act) else x
        return 'min %.1f  p50 %.1f  p90 %.1f  max %.1f us' % tuple(1e-3 * np.percentile(x, [0, 50, 90, 100]))
    print(f'{a.config} rows/gpu={kp.n_rows} grid={grid}x{warps} warps tiles={n_tiles} ({n_tiles / (grid * warps):.1f} per warp) '
          f'classes 1/2/3+ groups: {cls_total.tolist()}; tiles per warp min {per_warp.sum(1).min()} max {per_warp.sum(1).max()}')
    print(f'kernel span (first entry -> last exit): {1e-3 * span:.1f} us; spans of {a.evals} runs: '
          + ', '.join('%.1f' % (1e-3 * (r[:, 4].max() - r[:, 0].min())) for r in runs))
    print('entry after first entry :', q(t0 - base))
    print('prologue (entry->ready)  :', q(t1 - t0))
    if tr.shape[1] >= 12 and tr[:, 8].max() > 0:
        if tr.shape[1] >= 14 and tr[:, 12].max() > 0:
            print('    tables cleared (entry->):', q(tr[:, 13] - t0))
            print('    values written (->)     :', q(tr[:, 12] - tr[:, 13]))
            print('    first values requested, barrier (->):', q(tr[:, 8] - tr[:, 12]))
        print('  tables filled (entry->) :', q(tr[:, 8] - t0))
        print('  priors (->done)         :', q(tr[:, 9] - tr[:, 8]))
        print('  rest (->ready)          :', q(t1 - tr[:, 9]))
    print('wait for the first tile  :', q(t2 - t1))
    print('tile loop                :', q(t3 - t2))
    print('epilogue (loop end->exit):', q(t4 - t3))
    if tr.shape[1] >= 12 and tr[:, 10].max() > 0:
        print('  wait for the CTA (loop end->) :', q(tr[:, 10] - t3))
        print('  global partials (->stored)    :', q(tr[:, 11] - tr[:, 10]))
        print('  secondary partials (->exit)   :', q(t4 - tr[:, 11]))
    print('loop end after first entry:', q(t3 - base))
    # least-squares cost per tile class
    A = per_warp[act].astype(np.float64)
    y = (t3 - t2)[act].astype(np.float64) * 1e-3
    cols = [k for k in range(3) if A[:, k].any()]
    coef, *_ = np.linalg.lstsq(np.c_[A[:, cols], np.ones(len(A))], y, rcond=None)
    print('loop time ~ ' + ' + '.join(f'{c:.2f} us x tiles[{["1 group", "2 groups", "3+ groups"][k]}]' for k, c in zip(cols, coef))
          + f' + {coef[-1]:.2f} us')
    slow = np.argsort(-(t3 - base))[:5]
    for i in slow:
        print(f'  slowest warp cta {i // warps} warp {i % warps}: tiles {per_warp[i].tolist()} entry +{1e-3 * (t0[i] - base):.1f} '
              f'prologue {1e-3 * (t1[i] - t0[i]):.1f} wait {1e-3 * (t2[i] - t1[i]):.1f} loop {1e-3 * (t3[i] - t2[i]):.1f} '
              f'end +{1e-3 * (t3[i] - base):.1f} us')
    if len(fin):
        # the kernel behind the row kernel (finish / finish+exchange), per CTA class, relative to the row kernel
        last_exit = t4.max()
        names = ['primary slots', 'secondary slots', 'global + logp', 'hyper-parameters', 'unfed elements', 'multi-slot elements']
        print(f'finish kernel: {len(fin)} CTAs; first entry {1e-3 * (fin[:, 0].min() - base):+.1f} us after the row kernel\'s first entry, '
              f'{1e-3 * (fin[:, 0].min() - last_exit):+.1f} us relative to its last exit; last exit '
              f'{1e-3 * (fin[:, 3].max() - last_exit):+.1f} us after the row kernel\'s last exit')
        for c in range(6):
            sel = fin[fin[:, 4] == c]
            if not len(sel):
                continue
            rel = lambda col: 1e-3 * (sel[:, col] - last_exit)
            print(f'  {names[c]:20s} {len(sel):4d} CTAs: entry {rel(0).min():+6.1f}..{rel(0).max():+6.1f}  results visible {rel(1).min():+6.1f}..{rel(1).max():+6.1f}'
                  f'  own sums done {rel(2).min():+6.1f}..{rel(2).max():+6.1f}  exit {rel(3).min():+6.1f}..{rel(3).max():+6.1f} us')
        if fin.shape[1] >= 7 and fin[:, 5].max() > 0:
            slow = fin[np.argmax(fin[:, 3])]
            r = lambda col: 1e-3 * (int(slow[col]) - last_exit)
            print(f'  CTA that exits last (class {names[int(slow[4])]}): entry {r(0):+.1f}  results visible {r(1):+.1f}  own sums {r(2):+.1f}  '
                  f'slice combined {r(5):+.1f}  logp lines polled {r(6) if slow[6] else float("nan"):+.1f}  exit {r(3):+.1f} us')
    if a.out:
        json.dump({'config': a.config, 'rows': int(kp.n_rows), 'span_us': 1e-3 * float(span),
                   'trace': tr[:, :5].tolist(), 'tiles': per_warp.tolist()}, open(a.out, 'w'))
    f.close()


if __name__ == '__main__':
    main()
