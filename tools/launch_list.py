#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_list.py launches.csv > profiles/<name>.txt"""
import csv
import sys
from collections import OrderedDict


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    ig, ib = hdr.index('Grid Size'), hdr.index('Block Size')
    agg = OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(r[iu], 1.0)
        agg.setdefault((r[ik], r[ig], r[ib]), []).append(v)
    ours = sum(sum(v) / len(v) for (k, _, _), v in agg.items() if 'skk' in k)
    for (k, g, b), v in agg.items():
        share = f'share of evaluation {100 * (sum(v) / len(v)) / ours:5.1f}%' if 'skk' in k else ''
        print(f'{k[:70]:70s} grid {g:14s} block {b:12s} n={len(v):3d} avg {sum(v) / len(v):9.1f} us min {min(v):9.1f} us {share}')


if __name__ == '__main__':
    main(sys.argv[1])
