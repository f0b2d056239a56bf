#!/bin/bash
# tcgen05 X2H kernels: parity, then timing with CBG_EDGE_IMPL=6 (and the launch-level profile)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_x2h_tc.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1; echo "tc rc=$?"; tail -5 gpurun_out/pytest_tc.log | cut -c1-400
for impl in 6; do
CBG_EDGE_IMPL=$impl timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_impl$impl.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_impl$impl.log').read().strip().splitlines()[-1])
    print('impl $impl ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), {k:round(v['ms_per_step'],3) for k,v in (d.get('kernels') or {}).items() if v['ms_per_step']>0.03})
except Exception as e: print('parse fail', e); print(open('gpurun_out/bench_impl$impl.log').read()[-1500:])
PY
done
