#!/bin/bash
mkdir -p gpurun_out
for v in x1 x2 sc27; do ( ACB_LIB=$PWD/pyahocorasick_b200/_native/libacb200_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_batches or pathological or dense or pair_kernel" 2>&1 | tail -2 ); done
run() { name=$1; var=$2; cfg=$3; lib=$PWD/pyahocorasick_b200/_native/libacb200${name:+_$name}.so
  ACB_LIB=$lib timeout 200 python bench.py --config $cfg --steps 20 --warmup 5 --variant $var --no-cpu-baseline --no-e2e --no-latency 2>&1 | python tools/kline.py "lib=${name:-default} $cfg variant=$var"; }
( run "" planted C2; run x1 planted C2; run x2 planted C2; run x1 sparse C2; run sc27 planted C5; run "" planted C5; run sc27 planted C3 ) 2>&1 | tee gpurun_out/r2p_variants.log
