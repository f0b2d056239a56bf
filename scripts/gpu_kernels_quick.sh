#!/bin/bash
# edge-kernel iteration loop: kernel tests, parity suite (without the full-size file), bench, traces
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_x2h_tc.py tests/test_gpu_parity.py tests/test_f2_samplers.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_h2x.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/pytest_h2x.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_q.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_q.log').read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), {k:round(v['ms_per_step'],3) for k,v in (d.get('kernels') or {}).items() if v['ms_per_step']>0.03})
except Exception as e: print('parse fail', e); print(open('gpurun_out/bench_q.log').read()[-800:])
PY
timeout 200 python scripts/trace_node_gemm.py > gpurun_out/trace_node_gemm.txt 2>&1; echo "gemm trace rc=$?"; tail -11 gpurun_out/trace_node_gemm.txt | cut -c1-200
timeout 300 python scripts/trace_x2h_tc.py > gpurun_out/trace_x2h_tc.txt 2>&1; echo "trace rc=$?"; tail -14 gpurun_out/trace_x2h_tc.txt | cut -c1-200
