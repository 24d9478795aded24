#!/bin/bash
# GPU-side experiment helper: time the assembly pass with every variants/lib_*.so (built by
# tools/build_variant.sh) next to the shipped library.  Usage: tools/run_variants.sh [bench args]
ARGS="${@:---steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-solve --no-sweep}"
mkdir -p gpurun_out
echo "shipped: $(python bench.py $ARGS | tee gpurun_out/var_shipped.json | head -c 330)"
for so in variants/lib_*.so; do
  n=$(basename $so .so); n=${n#lib_}
  echo "$n: $(SIMUDO_B200_SO=$PWD/$so python bench.py $ARGS | tee gpurun_out/var_$n.json | head -c 330)"
done
