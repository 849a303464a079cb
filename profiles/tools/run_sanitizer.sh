#!/bin/bash
# compute-sanitizer over the kernels of libhast_b200.so: memcheck, racecheck (shared-memory hazards), initcheck,
# driven by the smoke test and a small slice of the parity tests.   usage: profiles/tools/run_sanitizer.sh TAG
TAG=${1:-rXX}; O=gpurun_out; mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
T='tests/test_gpu_parity.py::test_fused_parity tests/test_gpu_parity.py::test_long_reads_multi_pass tests/test_gpu_parity.py::test_table_under_pressure tests/test_gpu_parity.py::test_saturated_filter_queue_drains tests/test_stage03.py tests/test_stage00.py::test_count_table_matches_oracle tests/test_stage00.py::test_partitions_union_equals_whole'
for tool in memcheck racecheck initcheck synccheck; do
  timeout 600 $CS --tool $tool --error-exitcode 9 --log-file $O/${TAG}_sanitizer_$tool.log python -m pytest $T -m gpu -x -q -k "not direct and not tma or stage0" > $O/${TAG}_sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; tail -2 $O/${TAG}_sanitizer_${tool}_pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $O/${TAG}_sanitizer_$tool.log | tail -2
done
