#!/bin/bash
# compute-sanitizer evidence: memcheck, racecheck and synccheck on the tests that drive every kernel family
mkdir -p gpurun_out
SEL="tests/test_x2h_tc.py tests/test_ipa.py tests/test_gpu_parity.py::test_short_trajectory_matches_reference_golden tests/test_gpu_parity.py::test_receptive_field_pruning_is_exact tests/test_gpu_parity.py::test_cuda_graph_replay_is_bit_identical tests/test_f2_samplers.py::test_bp_trajectory_matches_golden_and_oracle tests/test_batch_builder.py"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python -m pytest $SEL -m gpu -q -x -p no:cacheprovider -k "not many_tiles" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitizer_$tool.log | head -8
done
