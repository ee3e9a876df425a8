#!/usr/bin/env python
"""
Hot spots of a kernel from an .ncu-rep captured with `--set full --import-source on` (needs `ncu` on PATH,
no GPU): warp-stall samples per SASS instruction, the top instructions with their dominant stall reason.

usage: python tools/ncu_hot_lines.py report.ncu-rep [top N] > profiles/<name>_hot.txt
"""
import csv
import io
import subprocess
import sys


def main(path, top=40):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source=sass'],
                         capture_output=True, text=True).stdout
    lines = raw.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.reader(io.StringIO('\n'.join(lines[start:]))))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith('stall_') and not h.endswith('(Not Issued)')]
    items = []
    total = 0
    for n, r in enumerate(rows[1:]):
        if len(r) < len(hdr):
            continue
        s = int(r[col['# Samples']] or 0)
        total += s
        stalls = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)
        items.append((s, n, r[col['Source']].strip(), stalls[:2], int(r[col['Instructions Executed']] or 0)))
    print(lines[0][:200])
    print(f'instructions: {len(items)}   samples: {total}')
    for s, n, src, stalls, execd in sorted(items, reverse=True)[:top]:
        why = ', '.join(f'{name} {cnt}' for cnt, name in stalls if cnt)
        print(f'{100.0 * s / max(total, 1):5.1f}%  #{n:5d}  exec {execd:8d}  {src[:70]:70s} {why}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
