#!/bin/bash
# Developer helper: quick nvcc build of the native kernels for one band count (default 3),
# to read ptxas register / spill numbers without the full 2-minute build.
#   tools/probe_build.sh [NB] [extra nvcc flags...]  ->  /tmp/probe.so, /tmp/probe.log
NB=${1:-3}; shift
cd "$(dirname "$0")/../simudo_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC \
     -Xptxas -v -DSB_PROBE_NB=$NB "$@" -o ${PROBE_OUT:-/tmp/probe.so} pdd_kernels.cu band_solver.cu > ${PROBE_LOG:-/tmp/probe.log} 2>&1
LOG=${PROBE_LOG:-/tmp/probe.log}; grep -E "error" $LOG | head
grep -A2 -E "k_rows_nat|k_moments_nat|k_generation_nat" $LOG | grep -E "Compiling|spill|Used" | paste - - - \
  | sed -e 's/ptxas info    ://g' -e "s/Compiling entry function '//" -e "s/' for 'sm_100a'//" | awk '{printf "%-55s stack %s spill_st %s spill_ld %s regs %s smem %s\n", $1, $2, $6, $10, $15, $(NF-2)}'
