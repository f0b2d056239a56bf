#!/bin/bash
# tcgen05 X2H kernels: parity, timing with CBG_EDGE_IMPL=6, ncu of the two kernels, new test files
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_x2h_tc.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1; echo "tc rc=$?"; tail -3 gpurun_out/pytest_tc.log | cut -c1-400
CBG_EDGE_IMPL=6 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_impl6.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_impl6.log').read().strip().splitlines()[-1])
    print('impl 6 ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), {k:round(v['ms_per_step'],3) for k,v in (d.get('kernels') or {}).items() if v['ms_per_step']>0.03})
except Exception as e: print('parse fail', e); print(open('gpurun_out/bench_impl6.log').read()[-1500:])
PY
CBG_EDGE_IMPL=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:x2h_tc_kernel -s 60 -c 2 -o gpurun_out/prof_x2h_tc -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --profile-steps 0 > gpurun_out/ncu_x2h_tc.log 2>&1; echo "ncu rc=$?"
timeout 900 python -m pytest tests/test_sample_driver.py tests/test_full_size_parity.py -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/pytest_new.log 2>&1; echo "new rc=$?"; tail -25 gpurun_out/pytest_new.log | cut -c1-300
