#!/bin/bash
# kernel-time variants without pytest + ncu captures of chosen variants
mkdir -p gpurun_out
tools/gpu_variants.sh "$@" > gpurun_out/r2c_variants.log 2>&1
i=0
for spec in $NCU_SPECS; do   # name:ENV=..:variant
  name=$(echo $spec | cut -d: -f1); envs=$(echo $spec | cut -d: -f2); var=$(echo $spec | cut -d: -f3)
  lib=$PWD/pyahocorasick_b200/_native/libacb200${name:+_$name}.so
  env $envs ACB_LIB=$lib timeout 600 ncu --set full --import-source on --clock-control none -k regex:acb_stream -s 3 -c 1 -f -o gpurun_out/r2c_$i python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-latency --variant $var > gpurun_out/r2c_ncu_$i.log 2>&1
  i=$((i+1))
done
cat gpurun_out/r2c_variants.log
