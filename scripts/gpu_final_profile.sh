#!/bin/bash
# correctness, full default bench, then ncu evidence of the same command (launch list + full-set captures)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log | cut -c1-300
timeout 300 python scripts/bench_f3.py > gpurun_out/bench_f3.log 2>&1; tail -1 gpurun_out/bench_f3.log | cut -c1-400
timeout 600 python scripts/bench_f2.py > gpurun_out/bench_f2.log 2>&1; tail -4 gpurun_out/bench_f2.log | cut -c1-300
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --profile-steps 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 430 -c 125 --csv \
    --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launches_final.log 2>&1
echo "launch list exit $?"
for K in x2h_k_mma2_kernel x2h_v_kernel node_gemm_ws_kernel h2x_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 \
      -f -o gpurun_out/prof3_$K $CMD > gpurun_out/ncu3_$K.log 2>&1
  echo "$K exit $?"
done
