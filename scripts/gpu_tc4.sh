#!/bin/bash
# tcgen05 X2H kernels + f16 node GEMM: parity, timing (CBG_EDGE_IMPL=6), whole GPU suite without the slow full-size tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_x2h_tc.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1; echo "tc rc=$?"; tail -3 gpurun_out/pytest_tc.log | cut -c1-400
for ng in f16 tf32; do
CBG_NODE_GEMM=$ng CBG_EDGE_IMPL=6 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_impl6_$ng.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_impl6_$ng.log').read().strip().splitlines()[-1])
    print('impl 6 gemm $ng ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), {k:round(v['ms_per_step'],3) for k,v in (d.get('kernels') or {}).items() if v['ms_per_step']>0.03})
except Exception as e: print('parse fail', e); print(open('gpurun_out/bench_impl6_$ng.log').read()[-1500:])
PY
done
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider --deselect tests/test_x2h_tc.py --deselect tests/test_full_size_parity.py > gpurun_out/pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
