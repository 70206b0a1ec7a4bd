#!/usr/bin/env python
"""Summarise an ncu report: per-launch key metrics (raw page) and the top stall sites (source page).
usage: python tools/ncu_summary.py report.ncu-rep [kernel-regex] [top-n]"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed.sum', 'sm__cycles_elapsed.max',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'l1tex__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_xu.sum', 'launch__grid_size',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']


def run(args):
    return subprocess.run(['ncu'] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else None
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx['Kernel Name']]
        if pat and pat not in name:
            continue
        print('==', name[:80])
        for k in KEYS:
            if k in idx:
                print(f'   {k:70s} {r[idx[k]]:>16s} {units[idx[k]]}')
    if not pat:
        return
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat]))))
    h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[h]
    idx = {x: i for i, x in enumerate(hdr)}
    data = []
    for r in rows[h + 1:]:
        if len(r) != len(hdr):
            break                      # next launch of the same kernel
        data.append(r)
    samp = lambda r: int(r[idx['# Samples']] or 0)
    tot = sum(samp(r) for r in data)
    stalls = [x for x in hdr if x.startswith('stall_') and 'Not Issued' not in x]
    agg = {s: sum(int(r[idx[s]] or 0) for r in data) for s in stalls}
    print('-- samples', tot, 'SASS instructions', len(data), 'executed (warp)', sum(int(r[idx['Instructions Executed']] or 0) for r in data))
    print('-- stall mix:', ', '.join(f'{k[6:]} {100 * v / max(tot, 1):.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(data, key=lambda r: -samp(r))[:topn]:
        st = sorted(((s, int(r[idx[s]] or 0)) for s in stalls), key=lambda kv: -kv[1])[:2]
        print(f'   {100 * samp(r) / max(tot, 1):5.1f}%  {r[idx["Source"]][:70]:70s} {st[0][0][6:]}={st[0][1]} {st[1][0][6:]}={st[1][1]}')


if __name__ == '__main__':
    main()
