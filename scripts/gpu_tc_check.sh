#!/bin/bash
mkdir -p gpurun_out
echo "== node projection tests (simt + tcgen05)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k node_projections -p no:cacheprovider 2>&1 | tail -25
echo "== full GPU suite with tcgen05 node GEMM (default)"
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu_tc.log 2>&1; tail -15 gpurun_out/pytest_gpu_tc.log
echo "== bench tc"
timeout 600 python bench.py --steps 10 --warmup 3 --e2e-steps 50 --no-cpu-baseline > gpurun_out/bench_tc.log 2>&1; tail -2 gpurun_out/bench_tc.log | cut -c1-400
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_tc.log').read().strip().splitlines()[-1]); print({k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
except Exception as e: print('parse fail', e)
PY
echo "== bench simt"
CBG_NODE_GEMM=simt timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_simt.log 2>&1; tail -1 gpurun_out/bench_simt.log | cut -c1-300
