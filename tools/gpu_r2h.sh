#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_batches or pathological or dense or golden" 2>&1 | tail -5 ) > gpurun_out/r2h_pytest.log
cat gpurun_out/r2h_pytest.log
run() { name=$1; var=$2; lib=$PWD/pyahocorasick_b200/_native/libacb200${name:+_$name}.so
  ACB_LIB=$lib timeout 120 python bench.py --steps 20 --warmup 5 --variant $var --no-cpu-baseline --no-e2e --no-latency 2>&1 | python tools/kline.py "lib=${name:-default} variant=$var"; }
( for spec in "$@"; do run ${spec%%:*} ${spec#*:}; done ) 2>&1 | tee gpurun_out/r2h_variants.log
