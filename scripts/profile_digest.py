#!/usr/bin/env python
"""Turn the artefacts of scripts/gpu_r2_validate.sh (gpurun_out/) into the tracked evidence under profiles/:
ncu summaries per kernel, the launch-share table of one bench step and profiles/ncu_traffic.json (DRAM bytes per launch
of the dominant kernels, read by bench.py for roofline.traffic).   python scripts/profile_digest.py <tag>"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')


def raw_metrics(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], rows[2]))


def dram_bytes(rep):
    """[(kernel name, dram bytes, duration us)] per captured launch of an .ncu-rep"""
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], dict(zip(rows[0], rows[1]))
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    res = []
    for vals in rows[2:]:
        m = dict(zip(hdr, vals))
        tot = sum(float(m[k].replace(',', '')) * scale.get(units[k], 1) for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        num = lambda k: float(m[k].replace(',', '')) if k in m and m[k] else None
        extra = {'lsu_shared_wavefronts': num('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'),
                 'tensor_pipe_active_pct': num('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
                 'issue_active_pct': num('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                 'sm_cycles_elapsed_max': num('sm__cycles_elapsed.max'), 'grid': num('launch__grid_size'),
                 'registers_per_thread': num('launch__registers_per_thread')}
        res.append((m['Kernel Name'], int(tot), float(m['gpu__time_duration.sum'].replace(',', '')), extra))
    return res


def main():
    """python scripts/profile_digest.py <tag> [round]   (round defaults to r02: reads gpurun_out/prof_<round>_*.ncu-rep,
    launches_<round>.csv, bench_full.json, pytest_gpu.log written by scripts/gpu_r2_validate.sh)"""
    tag = sys.argv[1] if len(sys.argv) > 1 else 'vX'
    rnd = sys.argv[2] if len(sys.argv) > 2 else 'r02'
    traffic = {'_source': f'ncu --set full --clock-control none, bench.py c2 shape (profiles/{rnd}_ncu_*_{tag}.txt)'}
    for fn in sorted(os.listdir(OUT)):
        if not (fn.startswith(f'prof_{rnd}_') and fn.endswith('.ncu-rep')):
            continue
        kernel = fn[len(f'prof_{rnd}_'):-len('.ncu-rep')]
        rep = os.path.join(OUT, fn)
        txt = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_summary.py'), rep], capture_output=True, text=True).stdout
        with open(os.path.join(PROF, f'{rnd}_ncu_{kernel}_{tag}.txt'), 'w') as f:
            f.write(txt.replace(OUT + os.sep, 'gpurun_out/'))
        if kernel == 'x2h_tc_kernel':       # launch 0 = attention weights (mode 0), launch 1 = aggregation (mode 1)
            for (name, nbytes, us, extra), fam in zip(dram_bytes(rep), ('x2h_k_tc', 'x2h_v_tc')):
                traffic[fam] = {'dram_bytes_per_launch': nbytes, 'duration_us': us, 'kernel': name, 'ncu': extra}
    if len(traffic) > 1:
        with open(os.path.join(PROF, 'ncu_traffic.json'), 'w') as f:
            json.dump(traffic, f, indent=1)
    lst = os.path.join(OUT, f'launches_{rnd}.csv')
    if os.path.exists(lst):
        shutil.copy(lst, os.path.join(PROF, f'{rnd}_launches_{tag}.csv'))
        rows = list(csv.reader(open(lst)))
        for i, r in enumerate(rows):
            if 'Kernel Name' in r:
                h, start = r, i + 1
                break
        ik, iv = h.index('Kernel Name'), h.index('Metric Value')
        cnt, tot = collections.Counter(), collections.Counter()
        for r in rows[start:]:
            if len(r) > iv:
                name = r[ik].replace('void ', '').replace('<unnamed>::', '').split('(EdgeArgs')[0].split('(NodeGemm')[0][:52]
                cnt[name] += 1
                tot[name] += float(r[iv].replace(',', '')) / 1000.0
        total = sum(tot.values())
        with open(os.path.join(PROF, f'{rnd}_launch_shares_{tag}.txt'), 'w') as f:
            f.write('# ncu launch list of `CBG_GRAPH=0 python bench.py --steps 3 --warmup 3 ...` (c2; cold-cache, serialised: compare SHARES)\n')
            f.write(f'total {total:.1f} us over {sum(cnt.values())} launches\n')
            for name, v in tot.most_common():
                f.write(f'{name:52s} n={cnt[name]:3d} total {v:9.1f} us  avg {v / cnt[name]:7.1f}  {100 * v / total:5.1f}%\n')
    for src, dst in (('bench_full.json', f'{rnd}_bench_{tag}_full.json'), ('pytest_gpu.log', f'{rnd}_pytest_gpu_{tag}.log'),
                     ('smoke.log', None), ('trace_node_gemm.txt', f'{rnd}_trace_node_gemm_{tag}.txt'),
                     ('trace_x2h_tc.txt', f'{rnd}_trace_x2h_tc_{tag}.txt')):
        p = os.path.join(OUT, src)
        if not os.path.exists(p):
            continue
        lines = open(p).read().strip().splitlines()
        if dst is None:          # smoke line goes to the end of the pytest log
            with open(os.path.join(PROF, f'{rnd}_pytest_gpu_{tag}.log'), 'a') as f:
                f.write(lines[-1] + '\n')
            continue
        with open(os.path.join(PROF, dst), 'w') as f:
            f.write((lines[-1] if dst.endswith('.json') else '\n'.join(lines[-40:])) + '\n')
    print('profiles updated for', rnd, tag)


if __name__ == '__main__':
    main()
