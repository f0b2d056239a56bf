#!/bin/bash
# programmatic dependent launch on/off: kernel tests + parity + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_x2h_tc.py tests/test_gpu_parity.py tests/test_f2_samplers.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_pdl.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/pytest_pdl.log | cut -c1-300
for pdl in 1 0; do
for w in c2 c1; do
CBG_PDL=$pdl timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 200 > gpurun_out/bench_pdl${pdl}_$w.log 2>&1
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_pdl${pdl}_$w.log').read().strip().splitlines()[-1])
    print('pdl=$pdl $w ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), 'e2e', d['e2e'] and round(d['e2e']['value'],3), {k:round(v['ms_per_step'],3) for k,v in (d.get('kernels') or {}).items() if v['ms_per_step']>0.05})
except Exception as e: print('parse fail', e); print(open('gpurun_out/bench_pdl${pdl}_$w.log').read()[-800:])
PY
done
done
