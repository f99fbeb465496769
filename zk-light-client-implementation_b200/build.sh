#!/bin/bash
# Build libb200plonky2.so in-tree for sm_100a.  Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
SRCS=$(ls csrc/*.cu)
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared \
     -o libb200plonky2.so $SRCS "$@"
