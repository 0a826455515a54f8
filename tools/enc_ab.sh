#!/bin/bash
# Run under gpurun: tools/enc_ab.sh <tag> [generic] -- encode parity tests, then timings of the grid kernel (and the generic one).
tag=${1:-ab}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_encode.py tests/test_gpu_random_sweep.py tests/test_host_encoder.py tests/test_tfrecord.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
tail -5 gpurun_out/${tag}_tests.log
VARIANTS=${VARIANTS:-0,1,2,auto} python tools/enc_sweep.py 1,16,64,256 2>&1 | tee gpurun_out/${tag}_grid.txt
if [ "$2" = "generic" ]; then
RONK_ENC_KERNEL=generic VARIANTS=auto python tools/enc_sweep.py 1,16,64,256 2>&1 | tee gpurun_out/${tag}_generic.txt
fi
ncu --set full --clock-control none --import-source on -k regex:match_encode_grid -s 2 -c 1 -o gpurun_out/${tag}_enc256 python tools/prof.py --stage encode --batch 256 --iters 3 > /dev/null 2>&1
