#!/bin/bash
# Run under gpurun: compute-sanitizer memcheck + racecheck over the round-2 kernels (grid encode kernel forced on small
# batches, rank kernel, record append, batched filter_min, backward kernels, tiered select / NMS).
mkdir -p gpurun_out
T='tests/test_gpu_encode.py::test_bboxes_encode_golden tests/test_gpu_encode.py::test_encode_many_gt_and_flags tests/test_gpu_encode.py::test_encode_ssd512_anchor_set tests/test_gpu_encode.py::test_encode_extreme_coordinates tests/test_gpu_postprocess.py::test_device_tpfp_state_matches_host_accumulation tests/test_gpu_postprocess.py::test_filter_min_pad_axis_safe_divide tests/test_loss_masks.py::test_localization_loss_and_smooth_l1_are_differentiable tests/test_host_encoder.py'
RONK_ENC_KERNEL=grid timeout 1500 compute-sanitizer --tool memcheck --launch-timeout 600 python -m pytest $T -m gpu -x -q > gpurun_out/san2_mem.log 2>&1
tail -4 gpurun_out/san2_mem.log
RONK_ENC_KERNEL=grid timeout 1500 compute-sanitizer --tool racecheck --launch-timeout 600 python -m pytest tests/test_gpu_encode.py::test_bboxes_encode_golden tests/test_gpu_encode.py::test_encode_many_gt_and_flags tests/test_gpu_postprocess.py::test_device_tpfp_state_matches_host_accumulation -m gpu -x -q > gpurun_out/san2_race.log 2>&1
tail -4 gpurun_out/san2_race.log
