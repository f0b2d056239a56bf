#!/bin/bash
# edge-kernel tests, then benches: "IMPL RCACHE" pairs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "edge_kernel or static_fast" --maxfail=5 -p no:cacheprovider > gpurun_out/pytest_edge.log 2>&1
tail -12 gpurun_out/pytest_edge.log
for cfg in "$@"; do
  set -- $cfg
  CBG_EDGE_IMPL=$1 CBG_RCACHE=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_i$1_rc$2.log 2>&1
  python - "$1" "$2" <<'PY'
import json, sys
f = f'gpurun_out/bench_i{sys.argv[1]}_rc{sys.argv[2]}.log'
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print('impl', sys.argv[1], 'rcache', sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'lig/s', round(d['value'], 3),
          {k: round(v['ms_per_step'], 3) for k, v in (d.get('kernels') or {}).items() if v['ms_per_step'] > 0.05})
except Exception as e:
    print('parse fail', f, e); print(open(f).read()[-800:])
PY
done
