#!/bin/bash
# H2X on the tcgen05 tile kernel + S1/PROD stall fixes: kernel tests, parity suite (without the full-size file), bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_x2h_tc.py tests/test_gpu_parity.py tests/test_f2_samplers.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_h2x.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/pytest_h2x.log | cut -c1-300
for impl in tc simt; do
CBG_H2X_IMPL=$impl timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_h2x_$impl.log 2>&1; echo "bench $impl rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_h2x_$impl.log').read().strip().splitlines()[-1])
    print('$impl ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), {k:round(v['ms_per_step'],3) for k,v in (d.get('kernels') or {}).items() if v['ms_per_step']>0.03})
except Exception as e: print('parse fail', e); print(open('gpurun_out/bench_h2x_$impl.log').read()[-800:])
PY
done
timeout 200 python bench.py --workload c1 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_h2x_c1.log 2>&1; tail -1 gpurun_out/bench_h2x_c1.log | cut -c1-200
timeout 300 python scripts/trace_x2h_tc.py > gpurun_out/trace_x2h_tc.txt 2>&1; echo "trace rc=$?"; tail -14 gpurun_out/trace_x2h_tc.txt | cut -c1-200
