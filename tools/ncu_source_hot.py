#!/usr/bin/env python
"""Hot source lines of one kernel in an .ncu-rep captured with --import-source on (read here, no GPU):
python tools/ncu_source_hot.py file.ncu-rep [kernel-id ':::1'] [top N]
Prints, per source line, warp-stall samples and executed warp instructions (the `cuda,sass` view of
`ncu --page source`), the per-file totals and the dominant stall reasons of the hottest lines."""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    kid = sys.argv[2] if len(sys.argv) > 2 else ':::1'
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass',
                          '--kernel-id', kid], capture_output=True, text=True).stdout
    path, hdr = None, None
    lines = {}  # (file, line) -> dict
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == 'File Path':
            path = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            continue
        if r[0] in ('Function Name',) or hdr is None or r[0] == '':
            continue
        try:
            d = dict(zip(hdr[4:], r[4:]))
            samples = int(d['# Samples'])
            inst = int(d['Instructions Executed'])
        except (ValueError, KeyError):
            continue
        stalls = {k: int(v) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v.isdigit()}
        lines[(path, int(r[0]))] = dict(src=r[1].strip(), samples=samples, inst=inst, stalls=stalls,
                                        thr=d.get('Avg. Threads Executed', ''))
    tot_s = sum(v['samples'] for v in lines.values()) or 1
    tot_i = sum(v['inst'] for v in lines.values()) or 1
    print(f'total samples {tot_s}, warp instructions {tot_i}')
    per_file = collections.Counter()
    per_file_i = collections.Counter()
    agg = collections.Counter()
    for (f, _), v in lines.items():
        per_file[f] += v['samples']
        per_file_i[f] += v['inst']
        for k, n in v['stalls'].items():
            agg[k] += n
    print('stall mix:', ', '.join(f'{k[6:]} {100 * n / tot_s:.0f}%' for k, n in agg.most_common(8)))
    for f, n in per_file.most_common():
        print(f'  {f:18s} samples {100 * n / tot_s:5.1f}%  inst {100 * per_file_i[f] / tot_i:5.1f}%')
    for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1]['samples'])[:top]:
        st = ', '.join(f'{k[6:]} {n}' for k, n in sorted(v['stalls'].items(), key=lambda kv: -kv[1])[:3] if n)
        print(f"{100 * v['samples'] / tot_s:5.1f}% s {100 * v['inst'] / tot_i:5.1f}% i  {f}:{ln:<4d} {v['src'][:70]:70s} [{st}]")


if __name__ == '__main__':
    main()
