#!/bin/bash
# Pj ring depth sweep (CBG_PJ_RING=<k><v>): per-kernel times of the default bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_x2h_tc.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_ring.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/pytest_ring.log
for r in 44 45 54 55 64 65; do
CBG_PJ_RING=$r timeout 200 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_ring_$r.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_ring_$r.log').read().strip().splitlines()[-1])
    print('ring $r ms/step', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in (d.get('kernels') or {}).items() if k in ('x2h_k','x2h_v','h2x')})
except Exception as e: print('parse fail', e); print(open('gpurun_out/bench_ring_$r.log').read()[-500:])
PY
done
