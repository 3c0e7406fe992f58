#!/bin/bash
# Development aid: build a second library with extra nvcc defines for ONE translation unit, for A/B runs on the GPU box.
#   tools/ab_build.sh tvf_tslab.cu -DPPO_TS_FAKE_CONST=1     ->  ab/libppo_base.so.bin, ab/libppo_var.so.bin
# On the box: cp ab/libppo_var.so.bin ppo_b200/csrc/libppo_b200.so  (sources unchanged, so the stamp still matches)
set -e
cd "$(dirname "$0")/.."
src=$1; shift
python -c "import __graft_entry__ as g; g.build()" > /dev/null
mkdir -p ab
cp ppo_b200/csrc/libppo_b200.so ab/libppo_base.so.bin
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -c -o /tmp/ab_var.o ppo_b200/csrc/$src
objs=$(ls ppo_b200/csrc/build/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o ab/libppo_var.so.bin $objs /tmp/ab_var.o
ls -la ab/
