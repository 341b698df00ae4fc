#!/bin/bash
# compare experimental builds (ACB_LIB) on the bench kernel time; usage: tools/gpu_variants.sh "name:ENV=.. ENV2=.." ...
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}; [ "$envs" = "$spec" ] && envs=""
  lib=$PWD/pyahocorasick_b200/_native/libacb200${name:+_$name}.so
  for var in planted sparse; do
    env $envs ACB_LIB=$lib timeout 120 python bench.py --steps 20 --warmup 5 --variant $var --no-cpu-baseline --no-e2e --no-latency 2>&1 | python tools/kline.py "lib=${name:-default} [$envs] variant=$var"
  done
done
