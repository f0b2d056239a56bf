#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 200 --no-cpu-baseline > gpurun_out/bench_n$N.log 2>&1
echo "exit $?"; tail -3 gpurun_out/bench_n$N.log | cut -c1-900
