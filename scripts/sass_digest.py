#!/usr/bin/env python
"""Per-kernel SASS digest of libcbg_b200.so: which instruction classes each kernel is built from (the evidence the
profiling guide asks for: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP/UTMALDG = TMA bulk copies,
LDGSTS = cp.async, HMMA = legacy mma.sync, FFMA2 = Blackwell packed fp32).   python scripts/sass_digest.py > profiles/<tag>"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'cbgbench_b200', 'libcbg_b200.so')
CLASSES = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'LDGSTS', 'SYNCS', 'HMMA', 'FFMA2', 'FMUL2', 'FADD2',
           'FFMA', 'MUFU', 'F2FP', 'SHFL', 'LDS', 'STS', 'LDG', 'STG', 'ATOM', 'RED', 'BAR', 'STL', 'LDL']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.split('\n'):
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if cur and m:
            op = m.group(1)
            kernels[cur]['_total'] += 1
            for c in CLASSES:
                if op == c or op.startswith(c + '.') or (c in ('UTCHMMA', 'HMMA') and op.startswith(c)):
                    kernels[cur][c] += 1
    demangle = subprocess.run(['c++filt'] + list(kernels), capture_output=True, text=True).stdout.split('\n')
    print(f'# SASS digest of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; static instruction counts per kernel)')
    for (name, cnt), dm in zip(kernels.items(), demangle):
        short = re.sub(r'\(anonymous namespace\)::', '', dm)
        short = re.sub(r'\(.*', '', short)
        cols = ' '.join(f'{c}={cnt[c]}' for c in CLASSES if cnt[c])
        print(f'{short:48s} total={cnt["_total"]:6d}  {cols}')


if __name__ == '__main__':
    main()
