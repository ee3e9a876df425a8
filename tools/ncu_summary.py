#!/usr/bin/env python
"""Text summary of an .ncu-rep (needs `ncu` on PATH, no GPU): key counters + stall breakdown.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_issued.avg.per_cycle_active', 'smsp__inst_executed.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.max', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
        print('kernel:', d.get('Kernel Name', ('?',))[0])
        for k in KEYS:
            if k in d:
                print(f'  {k:70s} {d[k][0]:>16s} {d[k][1]}')
        st = {h: float(v[0].replace(',', '')) for h, v in d.items()
              if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued') and v[0]}
        tot = sum(st.values()) or 1.0
        print('  warp stall samples:')
        for h, v in sorted(st.items(), key=lambda x: -x[1])[:10]:
            print(f'    {h.replace("smsp__pcsamp_warps_issue_stalled_", ""):40s} {100 * v / tot:5.1f}%')
        print()


if __name__ == '__main__':
    main(sys.argv[1])
