#!/bin/bash
# quick kernel-time matrix (diagnostic, not a bench value)
for inl in ${INLINE_LIST:-1 0}; do
  for v in planted sparse; do
    ACB_INLINE_RESOLVE=$inl timeout 120 python bench.py --steps 10 --warmup 3 --variant $v --no-cpu-baseline --no-e2e 2>&1 | python tools/kline.py "inline=$inl variant=$v"
  done
done
