#!/bin/bash
# Kernel experiments: build pycricodecs_b200/libcricodecs_b200_<name>.so with extra -D flags for hca_fast_kernels.cu only
# (one template variant: stereo, no joint tools), the other objects as built. Select it at run time with CRI_LIB_PATH.
# Usage: tools/build_variant.sh <name> <flags...>
set -e
NAME=$1; shift
HERE=$(cd "$(dirname "$0")/.." && pwd)
C=$HERE/pycricodecs_b200/csrc
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xcompiler -ffp-contract=off"
$NV -DCRI_DEV_ONE_VARIANT "$@" -c $C/hca_fast_kernels.cu -o $C/build/hca_fast_kernels_$NAME.o
OBJS=$(ls $C/build/*.o | grep -v "hca_fast_kernels")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $HERE/pycricodecs_b200/libcricodecs_b200_$NAME.so $OBJS $C/build/hca_fast_kernels_$NAME.o -lcudart_static -lpthread -ldl -lrt
echo built libcricodecs_b200_$NAME.so
