#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_x2h_tc.py -m gpu -q -p no:cacheprovider --durations=12 > gpurun_out/pytest_tc.log 2>&1; echo "tc rc=$?"; tail -18 gpurun_out/pytest_tc.log | cut -c1-200
CUDA_LAUNCH_BLOCKING=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --profile-steps 0 > gpurun_out/bench_blocking.log 2>&1; echo "blocking rc=$?"; grep -v CUDAEvent gpurun_out/bench_blocking.log | grep -B2 -A6 "Error\|error" | head -30 | cut -c1-300
timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python bench.py --workload c1 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --profile-steps 0 > gpurun_out/sanitizer_c1.log 2>&1; echo "sanitizer rc=$?"; grep -v CUDAEvent gpurun_out/sanitizer_c1.log | head -40 | cut -c1-300
