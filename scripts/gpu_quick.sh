#!/bin/bash
# correctness (full GPU suite) + one short bench with the default configuration
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), {k:round(v['ms_per_step'],3) for k,v in (d.get('kernels') or {}).items() if v['ms_per_step']>0.05})
except Exception as e: print('parse fail', e); print(open('gpurun_out/bench_quick.log').read()[-800:])
PY
