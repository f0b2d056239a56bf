#!/bin/bash
# A/B of the static fast path / dynamic scheduling: targeted tests + short benches
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "static_fast or edge_kernel or pruning or rcache" --maxfail=5 -p no:cacheprovider > gpurun_out/pytest_fast.log 2>&1
tail -3 gpurun_out/pytest_fast.log
for F in 1 0; do
  CBG_DYN_SCHED=$F timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_fast$F.log 2>&1
  python - "$F" <<'PY'
import json, sys
f = f'gpurun_out/bench_fast{sys.argv[1]}.log'
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print('dyn_sched', sys.argv[1], 'ms/step', round(d['ms_per_step'], 3), 'lig/s', round(d['value'], 3),
          {k: round(v['ms_per_step'], 3) for k, v in (d.get('kernels') or {}).items() if v['ms_per_step'] > 0.05})
except Exception as e:
    print('parse fail', f, e); print(open(f).read()[-800:])
PY
done
