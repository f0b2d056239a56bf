#!/bin/bash
# register / spill / instruction-mix digest of the x2h_tc kernels
cd "$(dirname "$0")/../cbgbench_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v -c x2h_tc.cu -o /tmp/x2h_tc.o 2>&1 | grep -E "error|fatal|spill|Used" | head
cuobjdump -sass /tmp/x2h_tc.o > /tmp/x2h_tc.sass
for fn in Lb0 Lb1; do
  awk -v f="$fn" '/Function :/{on = index($0, f) > 0} on' /tmp/x2h_tc.sass > /tmp/k.sass
  echo "x2h_tc_kernel<$fn>: instr=$(grep -c '/\*[0-9a-f]\{4\}\*/' /tmp/k.sass) maxR=$(grep -oE '\bR[0-9]+\b' /tmp/k.sass | sed 's/R//' | sort -n | tail -1)"
  grep -oE '\b(UTCHMMA|UTCBAR|LDTM|STTM|LDGSTS|UBLKCP|UTMALDG|F2FP|FFMA2|FMUL2|FADD2|HMMA|SHFL|LDS|STS|LDG|STG|STL|LDL|MUFU|SYNCS|BAR)\b' /tmp/k.sass | sort | uniq -c | tr '\n' ' '; echo
done
