#!/bin/bash
# Kernel experiments: build pycricodecs_b200/libcricodecs_b200_<name>.so with extra -D flags for ONE translation unit
# (default hca_fast_kernels.cu with one template variant; UNIT=hca_enc_kernels.cu etc. for another), the other objects as
# built. Select it at run time with CRI_LIB_PATH.   Usage: [UNIT=file.cu] tools/build_variant.sh <name> <flags...>
set -e
NAME=$1; shift
UNIT=${UNIT:-hca_fast_kernels.cu}
BASE=${UNIT%.cu}
HERE=$(cd "$(dirname "$0")/.." && pwd)
C=$HERE/pycricodecs_b200/csrc
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xcompiler -ffp-contract=off"
EXTRA=""
[ "$UNIT" = hca_fast_kernels.cu ] && EXTRA="-DCRI_DEV_ONE_VARIANT"
$NV $EXTRA "$@" -c $C/$UNIT -o $C/build/${BASE}__$NAME.o
OBJS=$(ls $C/build/*.o | grep -v "__" | grep -v "/${BASE}.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $HERE/pycricodecs_b200/libcricodecs_b200_$NAME.so $OBJS $C/build/${BASE}__$NAME.o -lcudart_static -lpthread -ldl -lrt
echo built libcricodecs_b200_$NAME.so
