#!/bin/bash
# ncu full-set capture of the tensor-core X2H kernels (one launch each, mid-network layer)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --profile-steps 0"
for K in x2h_k_mma_kernel x2h_v_mma_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 \
      -f -o gpurun_out/prof5_$K $CMD > gpurun_out/ncu5_$K.log 2>&1
  echo "$K exit $?"
done
