#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): python tools/ncu_summary.py file.ncu-rep [out.csv]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__pcsamp_sample_buffer_full']


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        lines.append(['kernel', r[hdr.index('Kernel Name')].split('(')[0], ''])
        for i, h in enumerate(hdr):
            if h in WANT or (h.startswith('smsp__average_warps_issue_stalled') and float(r[i] or 0) >= 0.3):
                lines.append([h, r[i], units[i]])
    w = csv.writer(open(sys.argv[2], 'w') if len(sys.argv) > 2 else sys.stdout)
    w.writerows(lines)


if __name__ == '__main__':
    main()
