#!/usr/bin/env python
"""Instructions executed and stall samples per CUDA source line of one kernel in an ncu report
(needs -lineinfo and --import-source on).  usage: python tools/ncu_lines.py report.ncu-rep kernel-regex [top-n]"""
import csv
import io
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda', '--kernel-name', 'regex:' + pat],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    files = {}
    cur = None
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1]
            continue
        if r[0] == 'Line No':
            hdr = r
            continue
        if r[0] == 'Function Name' or hdr is None or cur is None:
            continue
        if len(r) != len(hdr) or not r[0].isdigit():
            continue
        files.setdefault(cur, []).append(r)
        if False:
            break
    idx = {h: i for i, h in enumerate(hdr)}
    allrows = []
    seen = set()
    for f, rs in files.items():
        for r in rs:
            key = (f, r[0])
            if key in seen:          # later launches of the same kernel repeat the listing
                continue
            seen.add(key)
            allrows.append((int(r[idx['Instructions Executed']] or 0), int(r[idx['# Samples']] or 0), f.split('/')[-1], int(r[0]), r[1].strip()))
    ti = sum(a[0] for a in allrows)
    ts = sum(a[1] for a in allrows)
    print(f'total warp instructions {ti}, samples {ts}')
    for ins, smp, f, ln, src in sorted(allrows, key=lambda a: -a[0])[:topn]:
        print(f'{100 * ins / max(ti, 1):5.1f}% inst {100 * smp / max(ts, 1):5.1f}% smpl  {f}:{ln:<4d} {src[:110]}')


if __name__ == '__main__':
    main()
