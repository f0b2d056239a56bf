#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics, SASS opcode mix and
warp-stall sampling per kernel.  Usage: python scripts/ncu_summary.py <file.ncu-rep> [nodes]"""
import collections
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'sm__inst_executed_pipe_uniform.sum']


def ncu(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nodes = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    print(f'# {rep}')
    for vals in rows[2:]:
        print('==== ' + vals[hdr.index('Kernel Name')])
        for w in WANT:
            for i, h in enumerate(hdr):
                if h == w:
                    print(f'{w:70s} {vals[i]:>16s} {units[i]}')
    src = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'source', '--csv']))))
    blocks, cur = [], None
    for r in src:
        if len(r) >= 2 and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'hdr': None, 'rows': []}
            blocks.append(cur)
        elif cur is not None and len(r) > 2 and r[0] == 'Address':
            cur['hdr'] = r
        elif cur is not None and cur['hdr'] is not None and len(r) == len(cur['hdr']):
            cur['rows'].append(r)
    for b in blocks:
        h = b['hdr']
        if not h:
            continue
        iS, iE = h.index('Source'), h.index('Instructions Executed')
        stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
        ops, stalls, tot = collections.Counter(), collections.Counter(), 0
        for r in b['rows']:
            toks = r[iS].strip().split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
            op = op.split('.')[0]
            n = int(r[iE] or 0)
            ops[op] += n
            tot += n
            for i, name in stall_cols:
                stalls[name] += int(r[i] or 0)
        print('==== ' + b['name'])
        print(f'total warp instructions {tot}' + (f'  ({tot / nodes:.0f} per node)' if nodes else ''))
        print('opcode mix:', ', '.join(f'{o} {100 * n / max(tot, 1):.1f}%' for o, n in ops.most_common(16)))
        st = sum(stalls.values()) or 1
        print('stall samples:', ', '.join(f'{k[6:]} {100 * v / st:.1f}%' for k, v in stalls.most_common(9)))


if __name__ == '__main__':
    main()
