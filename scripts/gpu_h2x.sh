#!/bin/bash
# h2x pair kernel: targeted tests under a hard timeout (a barrier bug must not hang the box), then A/B benches
mkdir -p gpurun_out
CBG_H2X_PAIRS=${1:-4} timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_f2_samplers.py -m gpu -q -k "h2x_impl or golden or trajectory or bp" --maxfail=3 -p no:cacheprovider > gpurun_out/pytest_h2x.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_h2x.log
for cfg in "0 4" "1 4" "1 5"; do
  set -- $cfg
  CBG_H2X_IMPL=$1 CBG_H2X_PAIRS=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_h2x_$1_$2.log 2>&1
  python - "$1" "$2" <<'PY'
import json, sys
f = f'gpurun_out/bench_h2x_{sys.argv[1]}_{sys.argv[2]}.log'
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print('h2x impl', sys.argv[1], 'pairs', sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'lig/s', round(d['value'], 3),
          {k: round(v['ms_per_step'], 3) for k, v in (d.get('kernels') or {}).items() if v['ms_per_step'] > 0.05})
except Exception as e:
    print('parse fail', f, e); print(open(f).read()[-800:])
PY
done
