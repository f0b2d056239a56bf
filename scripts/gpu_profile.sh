#!/bin/bash
# ncu evidence for the bench command: (1) launch list with device times, (2) full-set captures of
# the dominant kernels.  Never a bench value: numbers printed under ncu are discarded.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --profile-steps 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 240 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
for K in x2h_k_kernel x2h_v_kernel h2x_kernel node_gemm_kernel edge_gate_kernel knn_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 12 -c 1 \
      -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "$K exit $?"
done
ls -la gpurun_out
