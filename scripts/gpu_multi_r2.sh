#!/bin/bash
# multi-GPU data path on hardware: usage gpu_multi_r2.sh N   (N = GPUs of this box)
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_sharding_nccl.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_nccl.log 2>&1; echo "nccl test rc=$?"; tail -5 gpurun_out/pytest_nccl.log | cut -c1-300
fi
run() {  # workload scaling extra
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $1 --scaling $2 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 200 $3 > gpurun_out/bench_${1}_${2}_n$N.log 2> gpurun_out/bench_${1}_${2}_n$N.err
  echo "bench $1 $2 n=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_${1}_${2}_n$N.log').read().strip().splitlines() if l.startswith('{')][-1])
    print('$1 $2 N=$N: ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), 'e2e', d['e2e'] and round(d['e2e']['value'],3))
    if d.get('ranks'): print('  limiting rank', d['ranks']['limiting_rank'], 'imbalance', round(d['ranks']['imbalance'],3), [ (r['graphs'], r['nodes'], round(r['ms_per_step'],3)) for r in d['ranks']['per_rank']])
except Exception as e: print('parse fail', e); print(open('gpurun_out/bench_${1}_${2}_n$N.err').read()[-1500:])
PY
}
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --workload c5 --scaling strong --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 200 > gpurun_out/bench_c5_strong_n1.log 2>&1; echo "c5 n1 rc=$?"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_c5_strong_n1.log').read().strip().splitlines() if l.startswith('{')][-1]); print('c5 strong N=1 ms/step', round(d['ms_per_step'],3), 'lig/s', round(d['value'],3), 'e2e', round(d['e2e']['value'],3))"
  timeout 600 python bench.py --workload c2 --scaling strong --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 200 > gpurun_out/bench_c2_strong_n1.log 2>&1
else
  run c5 strong
  run c2 strong
  run c2 weak
fi
