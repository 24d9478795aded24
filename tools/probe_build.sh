#!/bin/bash
# Developer helper: quick nvcc build of the native kernels for one band count (default 3),
# to read ptxas register / spill numbers without the full 2-minute build.
#   tools/probe_build.sh [NB] [extra nvcc flags...]  ->  /tmp/probe.so, /tmp/probe.log
NB=${1:-3}; shift
cd "$(dirname "$0")/../simudo_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC \
     -Xptxas -v -DSB_PROBE_NB=$NB "$@" -o /tmp/probe.so pdd_kernels.cu > /tmp/probe.log 2>&1
grep -E "error" /tmp/probe.log | head
grep -A2 -E "k_rows_nat|k_moments_nat|k_generation_nat" /tmp/probe.log | grep -E "Compiling|spill|Used" | paste - - - \
  | sed -e 's/ptxas info    ://g' -e "s/Compiling entry function '//" -e "s/' for 'sm_100a'//" | awk '{printf "%-55s stack %s spill_st %s spill_ld %s regs %s smem %s\n", $1, $2, $6, $10, $15, $(NF-2)}'
