#!/bin/bash
# quick kernel-time matrix (diagnostic, not a bench value)
for th in ${THREADS_LIST:-1024 896 768}; do
  for v in planted sparse; do
    ACB_THREADS=$th timeout 120 python bench.py --steps 10 --warmup 3 --variant $v --no-cpu-baseline --no-e2e 2>&1 | python tools/kline.py "threads=$th variant=$v"
  done
done
