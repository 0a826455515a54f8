#!/bin/bash
# tools/variant.sh <name> <source.cu> [-DFLAG ...]: link a copy of libronk.so whose <source.cu> is compiled with extra flags
# into ron_tensorflow_b200/_variants/<name>.so (git-ignored; travels with gpurun).  tools/post_ab.sh times each variant.
set -e
name=$1; src=$2; shift 2
here=$(cd "$(dirname "$0")/.." && pwd)
python -m ron_tensorflow_b200.build > /dev/null
obj=$here/ron_tensorflow_b200/csrc/_obj
tmp=$(mktemp -d)
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC "$@" -c -o $tmp/v.o $here/ron_tensorflow_b200/csrc/$src
objs=$(ls $obj/*.o | grep -v "/${src%.cu}.o")
mkdir -p $here/ron_tensorflow_b200/_variants
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $here/ron_tensorflow_b200/_variants/$name.so $objs $tmp/v.o
rm -rf $tmp
echo built _variants/$name.so
