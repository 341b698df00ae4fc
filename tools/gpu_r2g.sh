#!/bin/bash
# pair kernel: ring shapes (stages x slices per tile) -- parity of the riskiest, timings, one ncu capture of the default
mkdir -p gpurun_out
( ACB_LIB=$PWD/pyahocorasick_b200/_native/libacb200_s4x15.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_batches or pathological or dense or golden" 2>&1 | tail -5 ) > gpurun_out/r2g_pytest.log
cat gpurun_out/r2g_pytest.log
run() { name=$1; var=$2; lib=$PWD/pyahocorasick_b200/_native/libacb200${name:+_$name}.so
  ACB_LIB=$lib timeout 120 python bench.py --steps 20 --warmup 5 --variant $var --no-cpu-baseline --no-e2e --no-latency 2>&1 | python tools/kline.py "lib=${name:-default} variant=$var"; }
( run "" planted; run s4x15 planted; run s3x20 planted; run s4x15 sparse; run s3x20 sparse; run s4x15nosurv sparse ) 2>&1 | tee gpurun_out/r2g_variants.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:acb_pair -s 3 -c 1 -f -o gpurun_out/r2g_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-latency > gpurun_out/r2g_full.log 2>&1
ls -la gpurun_out/r2g_*
