mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_full.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), 'e2e', d['e2e'] and round(d['e2e']['value'],3), 'launches', d['gpu_launches'], 'frac', round(d['roofline']['frac'],3))
PY
