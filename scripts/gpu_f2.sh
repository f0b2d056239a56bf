#!/bin/bash
# row f2: GPU parity tests + the three-model step bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_f2.py --steps 20 --warmup 3 > gpurun_out/bench_f2.log 2> gpurun_out/bench_f2.err; cat gpurun_out/bench_f2.log; tail -5 gpurun_out/bench_f2.err
