#!/bin/bash
# quick kernel-time matrix (diagnostic, not a bench value): C2 planted / sparse
for v in planted sparse; do
  timeout 120 python bench.py --steps 20 --warmup 5 --variant $v --no-cpu-baseline --no-e2e 2>&1 | python tools/kline.py "variant=$v"
done
