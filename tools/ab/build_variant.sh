#!/bin/bash
# usage: tools/ab/build_variant.sh NAME [-DJP_...=...]...   -> tools/ab/libs/NAME.so  (A/B builds of the same library, JUSTPIC_LIB=...)
set -e
name=$1; shift
mkdir -p tools/ab/libs
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -diag-suppress 550,128 -Xcompiler -fPIC -shared "$@" \
  -Xptxas=-v -o tools/ab/libs/$name.so justpic/jl_b200/csrc/justpic_sm100a.cu 2> tools/ab/libs/$name.ptxas.log
grep -A2 "${KGREP:-k_move_gatherILi3}" tools/ab/libs/$name.ptxas.log | grep -E "stack|Used" | tr '\n' ' '; echo
