#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total and
mean duration, share of the captured time.  usage: python tools/launch_summary.py launches.csv [skip-first-N]"""
import csv
import re
import sys


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline='') as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1.0)
        name = re.sub(r'\(.*', '', r['Kernel Name'])
        name = name.split('::')[-1]
        rows.append((int(r['ID']), name, v))
    rows = [r for r in rows if r[0] >= skip]
    agg = {}
    for _, n, v in rows:
        c, t = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, t + v)
    tot = sum(t for _, t in agg.values())
    print(f'# {path}: {len(rows)} launches, {tot / 1e3:.3f} ms total (cold-cache, serialised: compare shares)')
    print(f'{"kernel":44s} {"launches":>8s} {"total_us":>12s} {"mean_us":>10s} {"share":>7s}')
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{n[:44]:44s} {c:8d} {t:12.1f} {t / c:10.2f} {100 * t / tot:6.2f}%')


if __name__ == '__main__':
    main()
