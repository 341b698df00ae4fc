#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_iter_long_set.py tests/test_gpu_parity.py -m gpu -x -q -k "iter_long or golden or pair_kernel" 2>&1 | tail -5 ) | tee gpurun_out/r2o_pytest.log
run() { name=$1; var=$2; cfg=$3; lib=$PWD/pyahocorasick_b200/_native/libacb200${name:+_$name}.so
  ACB_LIB=$lib timeout 200 python bench.py --config $cfg --steps 20 --warmup 5 --variant $var --no-cpu-baseline --no-e2e --no-latency 2>&1 | python tools/kline.py "lib=${name:-default} $cfg variant=$var"; }
( run "" planted C2; run relax planted C2; run relax sparse C2 ) 2>&1 | tee gpurun_out/r2o_variants.log
