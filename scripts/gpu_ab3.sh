#!/bin/bash
# A/B over an environment variable: gpu_ab3.sh VAR v1 v2 ...
mkdir -p gpurun_out
VAR=$1; shift
for V in "$@"; do
  env $VAR=$V timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_${VAR}_$V.log 2>&1
  python - "$VAR" "$V" <<'PY'
import json, sys
f = f'gpurun_out/bench_{sys.argv[1]}_{sys.argv[2]}.log'
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(sys.argv[1], sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'lig/s', round(d['value'], 3),
          {k: round(v['ms_per_step'], 3) for k, v in (d.get('kernels') or {}).items() if v['ms_per_step'] > 0.05})
except Exception as e:
    print('parse fail', f, e); print(open(f).read()[-800:])
PY
done
