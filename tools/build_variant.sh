#!/bin/sh
# tools/build_variant.sh NAME [extra nvcc flags...]  ->  variants/lib_NAME.so (git-ignored)
# Experiment helper: build the extension with extra -D flags next to the shipped one.
set -e
name=$1; shift
mkdir -p variants
cd "$(dirname "$0")/../simudo_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared \
  -Xcompiler -fPIC "$@" -o ../../variants/lib_$name.so pdd_kernels.cu band_solver.cu
