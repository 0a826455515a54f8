#!/bin/bash
# Run under gpurun: compute-sanitizer memcheck + racecheck over the kernels added late in round 2 (device AP, large
# top_k select, wide / staged NMS).  The staged / wide NMS kernels are forced onto the small NMS cases as well.
mkdir -p gpurun_out
T='tests/test_gpu_postprocess.py::test_crowded_81_classes_topk_10000 tests/test_gpu_postprocess.py::test_nms_batch_sort_clip_golden tests/test_gpu_postprocess.py::test_nms_random_vs_oracle tests/test_gpu_postprocess.py::test_average_precision_records_vs_oracle tests/test_gpu_postprocess.py::test_device_state_average_precision_matches_host'
for mode in "RONK_NMS_STAGED=1" "RONK_NMS_STAGED=0 RONK_NMS_WIDE=4"; do
  env $mode timeout 1200 compute-sanitizer --tool memcheck --launch-timeout 600 python -m pytest $T -m gpu -x -q > gpurun_out/san4_mem.log 2>&1
  echo "memcheck [$mode]: $(grep -E 'passed|failed|error' gpurun_out/san4_mem.log | tail -1) / $(grep 'ERROR SUMMARY' gpurun_out/san4_mem.log | tail -1)"
  env $mode timeout 1500 compute-sanitizer --tool racecheck --launch-timeout 600 python -m pytest $T -m gpu -x -q > gpurun_out/san4_race.log 2>&1
  echo "racecheck [$mode]: $(grep -E 'passed|failed|error' gpurun_out/san4_race.log | tail -1) / $(grep 'RACECHECK SUMMARY' gpurun_out/san4_race.log | tail -1)"
done
