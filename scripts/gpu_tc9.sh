#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_x2h_tc.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1; echo "tc rc=$?"; tail -3 gpurun_out/pytest_tc.log | cut -c1-400
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_impl6.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_impl6.log').read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), {k:round(v['ms_per_step'],3) for k,v in (d.get('kernels') or {}).items() if v['ms_per_step']>0.03})
except Exception as e: print('parse fail', e); print(open('gpurun_out/bench_impl6.log').read()[-600:])
PY
timeout 300 python scripts/trace_x2h_tc.py > gpurun_out/trace_x2h_tc.txt 2>&1; echo "trace rc=$?"; sed -n 2,9p gpurun_out/trace_x2h_tc.txt | cut -c1-200; tail -13 gpurun_out/trace_x2h_tc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:x2h_tc_kernel -s 60 -c 2 -o gpurun_out/prof_x2h_tc -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --profile-steps 0 > gpurun_out/ncu_x2h_tc.log 2>&1; echo "ncu rc=$?"
