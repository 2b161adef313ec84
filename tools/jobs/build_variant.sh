#!/bin/sh
# A/B build: recompile one source with extra flags and link it with the other objects into libbmm_b200_<name>.so
# usage: tools/jobs/build_variant.sh <name> <source.cu> <flags...>
set -e
NAME=$1; SRC=$2; shift 2
D=bmm_mcmc_b200
O=/tmp/bmm_variant_$NAME.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  --expt-relaxed-constexpr --expt-extended-lambda "$@" -c $D/csrc/$SRC -o $O
OBJS=$(ls $D/build/*.o | grep -v "/$(basename $SRC .cu).o")
/usr/local/cuda/bin/nvcc -shared -o $D/libbmm_b200_$NAME.so $OBJS $O -gencode arch=compute_100a,code=sm_100a -ldl -lpthread
echo built $D/libbmm_b200_$NAME.so
