#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/trace_x2h_tc.py > gpurun_out/trace_x2h_tc.txt 2>&1; echo "rc=$?"; cat gpurun_out/trace_x2h_tc.txt | cut -c1-200
