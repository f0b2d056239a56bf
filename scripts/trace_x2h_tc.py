#!/usr/bin/env python
"""Pipeline timeline of the tcgen05 attention-weight kernel (CTA 0) at the c2 shape: SM-clock stamps per tile.
Events: 0 S1 start (after ACC1 wait) | 1 G built | 2 pre loaded + Pj added | 3 pair barrier passed | 4 S1 done (hf0)
        5 S1 start (hf1) | 6 S1 done (hf1) | 7 EPI h0 ready | 8 EPI h1 ready | 9 EPI done
        10 issuer: G/Pi ready | 11 MMA1 issued | 12 issuer: A ready | 13 MMA2 issued | 14 PROD start | 15 PROD done"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from cbgbench_b200 import _lib, synthetic
from cbgbench_b200.targetdiff import TargetDiffB200

torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
L = _lib.lib()
_lib.check(L.cbg_set_edge_impl(6, 0))
T = 1000
model = TargetDiffB200(synthetic.targetdiff_config(num_steps=T))
model.load_state_dict(synthetic.seeded_state_dict(model, seed=0), strict=True)
model = model.to(dev).eval()
model.use_graph = False            # the one-shot trace hook applies to an eager launch, not to a replayed graph
batch = synthetic.make_batch([300] * 64, [24] * 64, seed=2024)
state = model.prepare(batch)
n_lig, K = state['n_lig'], model.num_classes
X = torch.empty((T + 1, n_lig, 3), device=dev)
Cc = torch.empty((T + 1, n_lig, K), device=dev)
X[T].copy_(state['x_lig'])
Cc[T].copy_(state['c_lig'])
model.run_steps(state, [999, 998, 997], X, Cc)
torch.cuda.synchronize()
NT = 40
buf = torch.zeros((NT + 1, 16), dtype=torch.int64, device=dev)      # row NT: kernel entry / end of prologue / exit of CTA 0
WHICH = -1 if (len(sys.argv) > 1 and sys.argv[1] == 'v') else 1        # 'v': the aggregation kernel
_lib.check(L.cbg_debug_x2h_trace(buf.data_ptr(), WHICH * NT))
model.run_steps(state, [996], X, Cc)          # one-shot hook: the first (layer 0) attention-weight launch of this step
torch.cuda.synchronize()
_lib.check(L.cbg_debug_x2h_trace(None, 0))
t = buf.cpu().numpy()
n = int((t[:NT, 0] > 0).sum())
base = t[0, 10] if t[0, 10] > 0 else t[:n].min()
print('kernel:', 'aggregation (MODE_V)' if WHICH < 0 else 'attention weights (MODE_K)', '- tiles recorded', n)
cta = t[NT]
if cta[0] > 0:
    last = t[n - 1]
    print(f'CTA 0: entry -> end of prologue {cta[1] - cta[0]} clks; prologue -> first MMA1 issued {t[0, 11] - cta[1]}; '
          f'entry -> S1 of tile 0 starts {t[0, 0] - cta[0]}; last S1 done -> exit {cta[2] - last[4]}; '
          f'last EPI done -> exit {cta[2] - last[9]}; whole kernel {cta[2] - cta[0]} clks for {n} tiles')
names = ['S1st', 'Gblt', 'preLd', 'pairB', 'S1dn', 'S1st1', 'S1dn1', 'EPIh0', 'EPIh1', 'EPIdn', 'GPrdy', 'MMA1i', 'Ardy', 'MMA2i', 'PRst', 'PRdn']
print('tile ' + ' '.join(f'{x:>7s}' for x in names))
for k in range(min(n, 16)):
    print(f'{k:4d} ' + ' '.join(f'{int(v - base):7d}' if v > 0 else '      -' for v in t[k]))
if n > 8:
    d = np.diff(t[4:n - 2, 0])
    print('steady-state period (S1 start to S1 start):', d.mean(), 'clks; per stage means:')
    s = t[4:n - 2]
    for a, b, lab in ((0, 1, 'G build'), (1, 2, 'TMEM ld + Pj'), (2, 3, 'sumsq + pair barrier'), (3, 4, 'normalise + split + st'),
                      (4, 0, 'S1 done -> next S1 start (wait ACC1)'), (10, 11, 'MMA1 issue'), (11, 12, 'issuer waits A ready'),
                      (12, 13, 'MMA2 issue (incl. ACC2FREE waits)'), (13, 10, 'issuer waits next G/Pi'), (7, 9, 'EPI h0 -> done'),
                      (14, 15, 'PROD tile')):
        if b == 0 or (a == 13 and b == 10):
            v = s[1:, b] - s[:-1, a]
        else:
            v = s[:, b] - s[:, a]
        print(f'  {lab:40s} {v.mean():8.0f}')
