#!/bin/bash
# full validation: GPU suite (all tests), smoke, default bench with e2e + reference legs, reference arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -22 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_full.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), 'e2e', d['e2e'] and round(d['e2e']['value'],3), 'launches', d['gpu_launches'])
    print('roofline', json.dumps(d['roofline'])[:900])
    print('cpu', json.dumps(d['cpu_baseline'])[:400]); print('ref_gpu', json.dumps(d.get('reference_gpu'))[:400])
    print({k:round(v['ms_per_step'],3) for k,v in (d.get('kernels') or {}).items() if v['ms_per_step']>0.03})
except Exception as e: print('parse fail', e); print(open('gpurun_out/bench_full.json').read()[-800:])
PY
timeout 900 python bench.py --impl reference-gpu --steps 20 --warmup 5 > gpurun_out/bench_refgpu.json 2> gpurun_out/bench_refgpu.err; echo "refgpu rc=$?"; cut -c1-700 gpurun_out/bench_refgpu.json; tail -c 300 gpurun_out/bench_refgpu.err
