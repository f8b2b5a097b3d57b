#!/bin/bash
# build a tuning variant of the library: tools/build_variant.sh NAME "-DRT_WAVE_MINB=4 ..."  -> tools/lib_NAME.so
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC $2 -shared \
    petershirleyraytracer_b200/csrc/rt_api.cu -o tools/lib_$1.so
