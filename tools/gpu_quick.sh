#!/bin/bash
# parity (dense cases first under the watchdog build) + kernel lines of all configs
mkdir -p gpurun_out
( ACB_LIB=$PWD/pyahocorasick_b200/_native/libacb200_wd.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_batches or pathological or dense" 2>&1 | tail -5 ) > gpurun_out/q_pytest.log
if grep -q "passed" gpurun_out/q_pytest.log && ! grep -q "failed" gpurun_out/q_pytest.log; then
  ( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) >> gpurun_out/q_pytest.log
fi
for c in C2 C3 C4 C5; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-e2e 2>&1 | python tools/kline.py "config=$c"
done > gpurun_out/q_lines.log
tail -8 gpurun_out/q_pytest.log; cat gpurun_out/q_lines.log
