#!/bin/bash
# Developer tool: build variants of the library that differ in compile-time switches of ONE translation unit, for A/B timing on one
# GPU box in one gpurun call.   usage: tools/ab_build.sh bm_fused.cu NAME "-DFLAG=.. -DFLAG2=.."   ->  u96_slam_b200/lib/ab/NAME.so
# Use with  U96_LIB=u96_slam_b200/lib/ab/NAME.so python tools/bm_time.py ...
set -e
cd "$(dirname "$0")/.."
python -c "from u96_slam_b200 import build; build.build()" > /dev/null
SRC=$1; NAME=$2; FLAGS=$3
mkdir -p u96_slam_b200/lib/ab
OBJ=u96_slam_b200/lib/ab/${NAME}_${SRC%.cu}.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $FLAGS -c u96_slam_b200/csrc/$SRC -o $OBJ
OBJS=$(ls u96_slam_b200/lib/obj/*.o | grep -v "/${SRC%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o u96_slam_b200/lib/ab/$NAME.so $OBJS $OBJ
rm -f $OBJ
echo u96_slam_b200/lib/ab/$NAME.so
