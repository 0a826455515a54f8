#!/bin/bash
# Run under gpurun: compute-sanitizer memcheck + racecheck over the smoke path and the small parity tests.
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_mem.log 2>&1
tail -3 gpurun_out/san_mem.log
timeout 900 compute-sanitizer --tool racecheck --launch-timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_race.log 2>&1
tail -3 gpurun_out/san_race.log
timeout 1500 compute-sanitizer --tool memcheck --launch-timeout 600 python -m pytest tests/test_ron_eval.py tests/test_gpu_postprocess.py -m gpu -x -q -k "golden or ron_eval or ssd512" > gpurun_out/san_mem_tests.log 2>&1
tail -4 gpurun_out/san_mem_tests.log
