#!/bin/bash
# Experiment builds: tools/build_variant.sh NAME file.cu "-DX=1 -DY=2"  ->  variants/libNAME.so (that one source recompiled with the
# extra definitions, linked with the objects of the regular build).  tools/variant_env.py loads such a library instead of the
# product one.  variants/ is git-ignored (*.so) and travels to the GPU box.
set -e
NAME=$1; SRC=$2; DEFS=$3
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/variants
OBJ=$ROOT/variants/${NAME}_$(basename $SRC .cu).o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 $DEFS -Xptxas -v \
     -c $ROOT/pynufft_b200/csrc/$SRC -o $OBJ 2> $ROOT/variants/${NAME}.ptxas.log
OTHERS=$(ls $ROOT/pynufft_b200/build/*.o | grep -v "/$(basename $SRC .cu).o")
nvcc -shared -o $ROOT/variants/lib${NAME}.so $OBJ $OTHERS -lcufft -Xlinker -rpath -Xlinker /usr/local/cuda/lib64
grep -A2 "k_interp_col\|k_gridding_col" $ROOT/variants/${NAME}.ptxas.log | grep "Used\|spill" | tr '\n' ' '; echo
