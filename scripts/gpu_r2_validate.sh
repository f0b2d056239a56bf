#!/bin/bash
# round-2 validation of the current tree: full GPU suite, smoke, default bench (e2e + reference legs),
# ncu launch list of the bench command and one --set full capture per dominant kernel, node-GEMM trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_full.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), 'e2e', d['e2e'] and round(d['e2e']['value'],3), 'launches', d['gpu_launches'])
    print('roofline', json.dumps(d['roofline'])[:600])
    print({k:round(v['ms_per_step'],3) for k,v in (d.get('kernels') or {}).items() if v['ms_per_step']>0.03})
except Exception as e: print('parse fail', e); print(open('gpurun_out/bench_full.json').read()[-800:])
PY
timeout 300 python scripts/trace_x2h_tc.py > gpurun_out/trace_x2h_tc.txt 2>&1; echo "x2h trace rc=$?"; tail -13 gpurun_out/trace_x2h_tc.txt | cut -c1-160
timeout 200 python scripts/trace_node_gemm.py > gpurun_out/trace_node_gemm.txt 2>&1; echo "trace rc=$?"; tail -15 gpurun_out/trace_node_gemm.txt | cut -c1-200
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --profile-steps 0"
CBG_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv \
    --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
for w in c1 c3 c5; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 100 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "$w: $(cut -c1-160 gpurun_out/bench_$w.json | tail -1)"
done
for K in x2h_tc_kernel node_gemm_f16_kernel; do
  CBG_GRAPH=0 timeout 500 ncu --set full --clock-control none --import-source on -k regex:$K -s 8 -c 4 \
      -f -o gpurun_out/prof_r02_$K $CMD > gpurun_out/ncu_r02_$K.log 2>&1
  echo "$K exit $?"
done
