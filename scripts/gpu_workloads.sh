#!/bin/bash
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d['config']['workload'][:40], 'nodes', d['config']['nodes_per_gpu'], 'ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), {k:round(v['ms_per_step'],3) for k,v in (d.get('kernels') or {}).items() if v['ms_per_step']>0.05})
except Exception as e: print(sys.argv[1], 'parse fail', e); print(open(sys.argv[1]).read()[-1200:])
PY
}
for W in c3 c5 c1; do
  timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_$W.log 2>&1; summ gpurun_out/bench_$W.log
done
