#!/bin/bash
# A/B variant of ONE kernel file: tools/ab_build.sh <tag> <file.cu> [-DNAME=VALUE ...]  ->  digital-subband-video-1_b200/build/ab/libdsv1_b200_<tag>.so
# (the other objects are the current build's); run it with DSV1_B200_LIB=<that .so> python tools/ab_kernel.py
set -e
cd "$(dirname "$0")/../digital-subband-video-1_b200"
tag=$1; src=$2; shift 2
make -s -j8
mkdir -p build/ab
base=$(basename "$src" .cu)
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../include -Icsrc "$@" -c "csrc/$base.cu" -o "build/ab/${base}_$tag.o"
objs=$(ls build/*.o | grep -v "build/$base.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "build/ab/libdsv1_b200_$tag.so" $objs "build/ab/${base}_$tag.o" -lcudart
echo "build/ab/libdsv1_b200_$tag.so"
