#!/usr/bin/env python
"""Turn the artefacts of scripts/gpu_final_profile.sh (gpurun_out/) into the tracked evidence under profiles/:
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


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'vX'
    traffic = {'_source': f'ncu --set full --clock-control none, one launch each, bench.py c2 shape (profiles/r01_ncu_*_{tag}.txt)'}
    for kernel, fam in (('x2h_k_mma2_kernel', 'x2h_k'), ('x2h_v_kernel', 'x2h_v'), ('node_gemm_ws_kernel', None), ('h2x_kernel', None)):
        rep = os.path.join(OUT, f'prof3_{kernel}.ncu-rep')
        if not os.path.exists(rep):
            continue
        txt = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_summary.py'), rep], capture_output=True, text=True).stdout
        with open(os.path.join(PROF, f'r01_ncu_{kernel}_{tag}.txt'), 'w') as f:
            f.write(txt)
        if fam:
            m = raw_metrics(rep)
            scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
            rows = list(csv.reader(io.StringIO(out)))
            units = dict(zip(rows[0], rows[1]))
            tot = 0.0
            for key in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                tot += float(m[key].replace(',', '')) * scale.get(units[key], 1)
            traffic[fam] = {'dram_bytes_per_launch': int(tot), 'duration_us': float(m['gpu__time_duration.sum'].replace(',', '')),
                            'kernel': kernel}
    with open(os.path.join(PROF, 'ncu_traffic.json'), 'w') as f:
        json.dump(traffic, f, indent=1)
    lst = os.path.join(OUT, 'launches_final.csv')
    if os.path.exists(lst):
        shutil.copy(lst, os.path.join(PROF, f'r01_launches_{tag}.csv'))
        rows = list(csv.reader(open(lst)))
        for i, r in enumerate(rows):
            if 'Kernel Name' in r:
                h, start = r, i + 1
                break
        ik, iv = h.index('Kernel Name'), h.index('Metric Value')
        cnt, tot = collections.Counter(), collections.Counter()
        for r in rows[start:]:
            if len(r) > iv:
                name = r[ik].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:44]
                cnt[name] += 1
                tot[name] += float(r[iv].replace(',', '')) / 1000.0
        total = sum(tot.values())
        with open(os.path.join(PROF, f'r01_launch_shares_{tag}.txt'), 'w') as f:
            f.write('# ncu launch list of `python bench.py --steps 3 --warmup 3 ...` (c2; cold-cache, serialised: compare SHARES)\n')
            f.write(f'total {total:.1f} us over {sum(cnt.values())} launches\n')
            for name, v in tot.most_common():
                f.write(f'{name:44s} n={cnt[name]:3d} total {v:9.1f} us  avg {v / cnt[name]:7.1f}  {100 * v / total:5.1f}%\n')
    for src, dst in (('bench_full.log', f'r01_bench_{tag}_full.json'), ('bench_f3.log', f'r01_bench_{tag}_f3.json'), ('bench_f2.log', f'r01_bench_{tag}_f2.jsonl'),
                     ('pytest_gpu.log', f'r01_pytest_gpu_{tag}.log')):
        p = os.path.join(OUT, src)
        if os.path.exists(p):
            lines = open(p).read().strip().splitlines()
            with open(os.path.join(PROF, dst), 'w') as f:
                f.write((lines[-1] if dst.endswith('.json') else '\n'.join(l for l in lines[-6:] if not dst.endswith('.jsonl') or l.startswith('{'))) + '\n')
    print('profiles updated for', tag)


if __name__ == '__main__':
    main()
