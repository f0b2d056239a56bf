#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --profile-steps 0"
for K in x2h_k_kernel x2h_v_kernel node_gemm_tc_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 12 -c 1 \
      -f -o gpurun_out/prof2_$K $CMD > gpurun_out/ncu2_$K.log 2>&1
  echo "$K exit $?"
done
