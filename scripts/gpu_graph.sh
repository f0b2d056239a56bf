#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "graph or trajectory or rng or pruning" > gpurun_out/pytest_graph.log 2>&1; echo "graph tests rc=$?"; tail -12 gpurun_out/pytest_graph.log | cut -c1-300
for g in 1 0; do
for w in c2 c1; do
CBG_GRAPH=$g timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 200 > gpurun_out/bench_graph${g}_$w.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_graph${g}_$w.log').read().strip().splitlines()[-1])
    print('graph=$g $w ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), 'e2e', d['e2e'] and round(d['e2e']['value'],3), 'launches', d['gpu_launches'])
except Exception as e: print('parse fail', e); print(open('gpurun_out/bench_graph${g}_$w.log').read()[-1200:])
PY
done
done
