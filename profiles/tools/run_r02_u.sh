#!/bin/bash
# round 2: compute-sanitizer over the fused kernel with the per-batch tile size (tile-size invariance test + fused parity + long reads),
# then the stage-script test after its date filter changed
O=gpurun_out; TAG=${1:-r02_u}; mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
T='tests/test_gpu_parity.py::test_tile_size_does_not_change_counts tests/test_gpu_parity.py::test_fused_parity tests/test_gpu_parity.py::test_long_reads_multi_pass tests/test_gpu_parity.py::test_very_long_reads_between_short_ones tests/test_gpu_parity.py::test_saturated_filter_queue_drains'
for tool in memcheck racecheck initcheck synccheck; do
  timeout 500 $CS --tool $tool --error-exitcode 9 --log-file $O/${TAG}_sanitizer_$tool.log python -m pytest $T -m gpu -x -q -k "not direct and not tma" > $O/${TAG}_sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; tail -1 $O/${TAG}_sanitizer_${tool}_pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $O/${TAG}_sanitizer_$tool.log | tail -1
done
python -m pytest tests/test_stage_script_gpu.py -m gpu -q 2>&1 | tail -1
