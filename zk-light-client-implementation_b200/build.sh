#!/bin/bash
# Build libb200plonky2.so in-tree for sm_100a (see Makefile).  Usage: ./build.sh [EXTRA="extra nvcc flags"]
set -e
cd "$(dirname "$0")"
make -j"$(nproc)" "$@"
