#!/bin/bash
# kernel-time variants only (no parity, no ncu); planted variant only
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}; [ "$envs" = "$spec" ] && envs=""
  lib=$PWD/pyahocorasick_b200/_native/libacb200${name:+_$name}.so
  env $envs ACB_LIB=$lib timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-latency 2>&1 | python tools/kline.py "lib=${name:-default} [$envs]"
done | tee gpurun_out/var_only.log
