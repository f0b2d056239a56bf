#!/bin/bash
# tests, GEMM ring comparison, full default bench (e2e 1000 steps + cpu baseline), reference arm, 2-GPU bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), 'e2e', (d.get('e2e') or {}).get('value'), {k:round(v['ms_per_step'],3) for k,v in (d.get('kernels') or {}).items() if v['ms_per_step']>0.05})
except Exception as e: print(sys.argv[1], 'parse fail', e); print(open(sys.argv[1]).read()[-800:])
PY
}
for S in 2 3; do
  CBG_OVERLAP=0 CBG_GEMM_STAGES=$S timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_st${S}_noov.log 2>&1; summ gpurun_out/bench_st${S}_noov.log
done
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; summ gpurun_out/bench_full.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-600
