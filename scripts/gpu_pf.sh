#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), {k:round(v['ms_per_step'],3) for k,v in (d.get('kernels') or {}).items() if v['ms_per_step']>0.05})
except Exception as e: print(sys.argv[1], 'parse fail', e); print(open(sys.argv[1]).read()[-800:])
PY
}
for PF in 1 2 0; do
CBG_EDGE_PF=$PF timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_pf$PF.log 2>&1; summ gpurun_out/bench_pf$PF.log
done
