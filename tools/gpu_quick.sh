#!/bin/bash
# parity (dense cases first under the watchdog build), then the bench line
mkdir -p gpurun_out
( ACB_LIB=$PWD/pyahocorasick_b200/_native/libacb200_wd.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_batches or pathological or dense" 2>&1 | tail -5 ) > gpurun_out/q_pytest.log
if grep -q "passed" gpurun_out/q_pytest.log && ! grep -q "failed" gpurun_out/q_pytest.log; then
  ( timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 ) >> gpurun_out/q_pytest.log
fi
timeout 600 python bench.py > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
tail -30 gpurun_out/q_pytest.log; python tools/kline.py "bench C2" < gpurun_out/q_bench.json; tail -3 gpurun_out/q_bench.err
