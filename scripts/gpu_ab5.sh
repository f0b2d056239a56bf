#!/bin/bash
# h2x pair kernel: targeted tests (under a hard timeout: a barrier bug must not hang the box), then A/B benches
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "h2x_impl or golden or trajectory" --maxfail=3 -p no:cacheprovider > gpurun_out/pytest_h2x.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_h2x.log
bash scripts/gpu_ab3.sh CBG_H2X_IMPL 1 0
